/* oracle/oracle_wt.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See oracle.h.
 *
 * Part 2: wt_huff<> (row a7) and wt_int<> (row a8): construction, rank / select / access /
 * inverse_select, and the reference's serialised form.
 * Citations are relative to /root/reference/include/sdsl/.
 */
#include "oracle_priv.h"

#include <stdlib.h>
#include <string.h>

#define UNDEF16 0xFFFFu

/* ------------------------------------------------------------------------------------------ */
/* a7: wt_huff<> = wt_pc<huff_shape, bit_vector, rank_support_v<1>, select_support_mcl<1>,     */
/*                       select_support_mcl<0>, byte_tree<>>   (wt_huff.hpp:62-67)            */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    uint64_t freq, sym, parent, child[2];
} pc_node; /* wt_helper.hpp:68-84 */

/* wt_huff.hpp:82-115: min-heap of (freq, node nr), ties by node nr; first popped = child 0 */
static uint64_t huff_shape(const uint64_t *C, uint64_t csize, pc_node *t)
{
    uint64_t n = 0, c, alive_cnt;
    uint8_t *alive;
    for (c = 0; c < csize; ++c)
        if (C[c] > 0) {
            t[n].freq = C[c];
            t[n].sym = c;
            t[n].parent = t[n].child[0] = t[n].child[1] = ~0ULL;
            ++n;
        }
    alive = (uint8_t *)calloc(2 * n + 1, 1);
    for (c = 0; c < n; ++c)
        alive[c] = 1;
    alive_cnt = n;
    while (alive_cnt > 1) {
        uint64_t v[2], k, j;
        for (k = 0; k < 2; ++k) { /* pop the smallest (freq, index) pair */
            uint64_t best = ~0ULL;
            for (j = 0; j < n; ++j)
                if (alive[j] && (best == ~0ULL || t[j].freq < t[best].freq))
                    best = j; /* strict '<' keeps the smaller index on ties */
            alive[best] = 0;
            v[k] = best;
        }
        t[v[0]].parent = n;
        t[v[1]].parent = n;
        t[n].freq = t[v[0]].freq + t[v[1]].freq;
        t[n].sym = 0;
        t[n].parent = ~0ULL;
        t[n].child[0] = v[0];
        t[n].child[1] = v[1];
        alive[n] = 1;
        ++n;
        --alive_cnt;
    }
    free(alive);
    return n;
}

/* wt_helper.hpp:230-317: BFS relabel, bv_pos, c_to_leaf, paths */
static uint64_t byte_tree_build(orc_wt_huff *w, const pc_node *t, uint64_t nt)
{
    uint64_t bv_size = 0, node_cnt = 1, head = 0, tail = 0, c, prev_c = 0;
    uint16_t *q = (uint16_t *)malloc(sizeof(uint16_t) * (nt + 1));
    w->nnodes = nt;
    w->nodes = (orc_wt_node *)calloc(nt, sizeof(orc_wt_node));
#define FROM_PC(dst, src)                                                                                              \
    do {                                                                                                               \
        (dst).bv_pos = (src).freq;                                                                                     \
        (dst).bv_pos_rank = (src).sym;                                                                                 \
        (dst).parent = (uint16_t)(src).parent;                                                                         \
        (dst).child[0] = (uint16_t)(src).child[0];                                                                     \
        (dst).child[1] = (uint16_t)(src).child[1];                                                                     \
    } while (0)
    FROM_PC(w->nodes[0], t[nt - 1]);
    q[tail++] = 0;
    while (head < tail) {
        uint16_t idx = q[head++];
        uint64_t frq = w->nodes[idx].bv_pos;
        int k;
        w->nodes[idx].bv_pos = bv_size;
        if (w->nodes[idx].child[0] != UNDEF16)
            bv_size += frq;
        if (w->nodes[idx].child[0] != UNDEF16)
            for (k = 0; k < 2; ++k) {
                FROM_PC(w->nodes[node_cnt], t[w->nodes[idx].child[k]]);
                w->nodes[node_cnt].parent = idx;
                q[tail++] = (uint16_t)node_cnt;
                w->nodes[idx].child[k] = (uint16_t)node_cnt++;
            }
    }
    free(q);
    for (c = 0; c < 256; ++c)
        w->c_to_leaf[c] = UNDEF16;
    for (c = 0; c < nt; ++c)
        if (w->nodes[c].child[0] == UNDEF16)
            w->c_to_leaf[(uint8_t)w->nodes[c].bv_pos_rank] = (uint16_t)c;
    for (c = 0; c < 256; ++c) {
        if (w->c_to_leaf[c] != UNDEF16) {
            uint16_t v = w->c_to_leaf[c];
            uint64_t pw = 0, pl = 0;
            while (v != 0) {
                pw <<= 1;
                if (w->nodes[w->nodes[v].parent].child[1] == v)
                    pw |= 1;
                ++pl;
                v = w->nodes[v].parent;
            }
            w->path[c] = pw | (pl << 56);
            prev_c = c;
        } else {
            w->path[c] = prev_c; /* length 0; the reference stores prev_c in the low bits (:313-315) */
        }
    }
    return bv_size;
}

orc_wt_huff *orc_wt_huff_build(const uint8_t *text, uint64_t n)
{
    orc_wt_huff *w = (orc_wt_huff *)calloc(1, sizeof(*w));
    uint64_t C[256], csize = 0, k, nt, bits, *node_pos;
    pc_node t[512];
    w->size = n;
    if (n == 0) /* wt_pc.hpp:196-197: everything stays default-constructed */
        return w;
    memset(C, 0, sizeof(C));
    for (k = 0; k < n; ++k) { /* wt_helper.hpp:42-55 */
        if ((uint64_t)text[k] >= csize)
            csize = (uint64_t)text[k] + 1;
        ++C[text[k]];
    }
    for (k = 0; k < csize; ++k)
        w->sigma += C[k] > 0; /* wt_helper.hpp:57-66 */
    nt = huff_shape(C, csize, t);
    bits = byte_tree_build(w, t, nt);
    w->bv_bits = bits;
    w->bv = (uint64_t *)calloc(((bits + 63) >> 6) + 2, 8);
    node_pos = (uint64_t *)malloc(8 * nt);
    for (k = 0; k < nt; ++k)
        node_pos[k] = w->nodes[k].bv_pos;
    for (k = 0; k < n; ++k) { /* wt_pc.hpp:97-111,218-240: one bit per level along the symbol's path */
        uint64_t p = w->path[text[k]];
        uint32_t len = (uint32_t)(p >> 56), l;
        uint16_t v = 0;
        for (l = 0; l < len; ++l, p >>= 1) {
            if (p & 1)
                w->bv[node_pos[v] >> 6] |= 1ULL << (node_pos[v] & 63);
            ++node_pos[v];
            v = w->nodes[v].child[p & 1];
        }
    }
    free(node_pos);
    w->rank_table = (uint64_t *)calloc(orc_rank_v_table_words(bits), 8);
    orc_rank_v_build(w->bv, bits, 1, w->rank_table);
    w->sel1 = orc_select_mcl_build(w->bv, bits, 1);
    w->sel0 = orc_select_mcl_build(w->bv, bits, 0);
    for (k = 0; k < nt; ++k) /* wt_helper.hpp:319-327 */
        if (w->nodes[k].child[0] != UNDEF16)
            w->nodes[k].bv_pos_rank = orc_rank_v(w->bv, w->rank_table, 1, w->nodes[k].bv_pos);
    return w;
}

void orc_wt_huff_free(orc_wt_huff *w)
{
    if (!w)
        return;
    free(w->bv);
    free(w->rank_table);
    orc_select_mcl_free(w->sel1);
    orc_select_mcl_free(w->sel0);
    free(w->nodes);
    free(w);
}

static uint64_t wt_bv_rank(const orc_wt_huff *w, uint64_t i)
{
    return orc_rank_v(w->bv, w->rank_table, 1, i);
}

/* wt_pc.hpp:371-399 */
uint64_t orc_wt_huff_rank(const orc_wt_huff *w, uint64_t i, uint8_t c)
{
    uint64_t p, r = i;
    uint32_t len, l;
    uint16_t v = 0;
    if (w->size == 0 || w->c_to_leaf[c] == UNDEF16)
        return 0;
    if (w->sigma == 1)
        return i;
    p = w->path[c];
    len = (uint32_t)(p >> 56);
    for (l = 0; l < len && r; ++l, p >>= 1) {
        uint64_t o = wt_bv_rank(w, w->nodes[v].bv_pos + r) - w->nodes[v].bv_pos_rank;
        r = (p & 1) ? o : r - o;
        v = w->nodes[v].child[p & 1];
    }
    return r;
}

/* wt_pc.hpp:411-430 (inverse_select) and :336-357 (operator[]) */
uint64_t orc_wt_huff_inverse_select(const orc_wt_huff *w, uint64_t i, uint64_t *sym)
{
    uint16_t v = 0;
    while (w->nodes[v].child[0] != UNDEF16) {
        uint64_t pos = w->nodes[v].bv_pos + i;
        uint64_t o = wt_bv_rank(w, pos) - w->nodes[v].bv_pos_rank;
        int bit = (int)((w->bv[pos >> 6] >> (pos & 63)) & 1);
        i = bit ? o : i - o;
        v = w->nodes[v].child[bit];
    }
    *sym = w->nodes[v].bv_pos_rank;
    return i;
}

/* wt_pc.hpp:443-474 */
uint64_t orc_wt_huff_select(const orc_wt_huff *w, uint64_t i, uint8_t c)
{
    uint64_t p, r;
    uint32_t len, l;
    uint16_t v;
    if (w->size == 0)
        return 0;
    v = w->c_to_leaf[c];
    if (v == UNDEF16)
        return w->size;
    if (w->sigma == 1)
        return (i - 1 < w->size) ? i - 1 : w->size;
    r = i - 1;
    p = w->path[c];
    len = (uint32_t)(p >> 56);
    p <<= (64 - len);
    for (l = 0; l < len; ++l, p <<= 1) {
        v = w->nodes[v].parent;
        if ((p & 0x8000000000000000ULL) == 0)
            r = orc_select_mcl(w->sel0, w->bv, w->nodes[v].bv_pos - w->nodes[v].bv_pos_rank + r + 1) - w->nodes[v].bv_pos;
        else
            r = orc_select_mcl(w->sel1, w->bv, w->nodes[v].bv_pos_rank + r + 1) - w->nodes[v].bv_pos;
    }
    return r;
}

void orc_wt_huff_rank_batch(const orc_wt_huff *w, const uint64_t *i, const uint8_t *c, uint64_t n, uint64_t *out)
{
    uint64_t k;
    for (k = 0; k < n; ++k)
        out[k] = orc_wt_huff_rank(w, i[k], c[k]);
}
void orc_wt_huff_select_batch(const orc_wt_huff *w, const uint64_t *i, const uint8_t *c, uint64_t n, uint64_t *out)
{
    uint64_t k;
    for (k = 0; k < n; ++k)
        out[k] = orc_wt_huff_select(w, i[k], c[k]);
}
void orc_wt_huff_access_batch(const orc_wt_huff *w, const uint64_t *i, uint64_t n, uint64_t *sym, uint64_t *rnk)
{
    uint64_t k, r;
    for (k = 0; k < n; ++k) {
        r = orc_wt_huff_inverse_select(w, i[k], &sym[k]);
        if (rnk)
            rnk[k] = r;
    }
}

/* wt_pc.hpp:713-726 + wt_helper.hpp:362-375,139-150 */
void orc__wt_huff_serialize_into(orc_buf *b, const orc_wt_huff *w)
{
    uint64_t k;
    orc__buf_u64(b, w->size);
    orc__buf_u64(b, w->sigma);
    orc__bv_serialize_into(b, w->bv, w->bv_bits);
    orc__rank_v_serialize_into(b, w->rank_table, w->bv_bits);
    orc__select_mcl_serialize_into(b, w->sel1);
    orc__select_mcl_serialize_into(b, w->sel0);
    orc__buf_u64(b, w->nnodes);
    for (k = 0; k < w->nnodes; ++k) { /* 22 bytes per node */
        orc__buf_u64(b, w->nodes[k].bv_pos);
        orc__buf_u64(b, w->nodes[k].bv_pos_rank);
        orc__buf_put(b, &w->nodes[k].parent, 2);
        orc__buf_put(b, w->nodes[k].child, 4);
    }
    orc__buf_put(b, w->c_to_leaf, 512);
    orc__buf_put(b, w->path, 2048);
}
uint64_t orc_wt_huff_serialize(const orc_wt_huff *w, uint8_t *out, uint64_t cap)
{
    orc_buf b = {0, 0, 0};
    orc__wt_huff_serialize_into(&b, w);
    return orc__buf_finish(&b, out, cap);
}

/* ------------------------------------------------------------------------------------------ */
/* a8: wt_int<> = wt_int<bit_vector, rank_support_v<1>, select_support_mcl<1>, <0>>            */
/* ------------------------------------------------------------------------------------------ */

/* wt_int.hpp:160-260: level k holds, for the sequence stably sorted by its top k bits, bit (max_level-k-1) */
orc_wt_int *orc_wt_int_build(const uint64_t *seq, uint64_t n)
{
    orc_wt_int *w = (orc_wt_int *)calloc(1, sizeof(*w));
    uint64_t max_elem = 1, i, *cur, *nxt, bits;
    uint32_t k;
    w->size = n;
    if (n == 0)
        return w;
    for (i = 0; i < n; ++i)
        if (seq[i] > max_elem)
            max_elem = seq[i];
    w->max_level = orc_hi(max_elem) + 1;
    bits = n * w->max_level;
    w->tree_bits = bits;
    w->tree = (uint64_t *)calloc(((bits + 63) >> 6) + 2, 8);
    cur = (uint64_t *)malloc(8 * n);
    nxt = (uint64_t *)malloc(8 * n);
    memcpy(cur, seq, 8 * n);
    for (k = 0; k < w->max_level; ++k) {
        uint32_t shift = w->max_level - k - 1;
        uint64_t start = 0;
        while (start < n) {
            uint64_t node = (shift + 1 >= 64) ? 0 : (cur[start] >> (shift + 1)), end = start, c0 = 0, c1 = 0, z;
            while (end < n && ((shift + 1 >= 64) ? 0 : (cur[end] >> (shift + 1))) == node)
                ++end;
            for (i = start; i < end; ++i)
                c0 += !((cur[i] >> shift) & 1);
            z = 0;
            for (i = start; i < end; ++i) {
                uint64_t pos = (uint64_t)k * n + i;
                if ((cur[i] >> shift) & 1) {
                    w->tree[pos >> 6] |= 1ULL << (pos & 63);
                    nxt[start + c0 + c1++] = cur[i];
                } else
                    nxt[start + z++] = cur[i];
            }
            if (k + 1 == w->max_level)
                w->sigma += (c0 > 0) + (c1 > 0);
            start = end;
        }
        {
            uint64_t *t = cur;
            cur = nxt;
            nxt = t;
        }
    }
    free(cur);
    free(nxt);
    w->rank_table = (uint64_t *)calloc(orc_rank_v_table_words(bits), 8);
    orc_rank_v_build(w->tree, bits, 1, w->rank_table);
    w->sel1 = orc_select_mcl_build(w->tree, bits, 1);
    w->sel0 = orc_select_mcl_build(w->tree, bits, 0);
    return w;
}

void orc_wt_int_free(orc_wt_int *w)
{
    if (!w)
        return;
    free(w->tree);
    free(w->rank_table);
    orc_select_mcl_free(w->sel1);
    orc_select_mcl_free(w->sel0);
    free(w);
}

static uint64_t ti_rank(const orc_wt_int *w, uint64_t i)
{
    return orc_rank_v(w->tree, w->rank_table, 1, i);
}

/* wt_int.hpp:379-409 */
uint64_t orc_wt_int_rank(const orc_wt_int *w, uint64_t i, uint64_t c)
{
    uint64_t offset = 0, mask, node_size = w->size;
    uint32_t k;
    if (w->size == 0 || (w->max_level < 64 && (1ULL << w->max_level) <= c))
        return 0;
    mask = 1ULL << (w->max_level - 1);
    for (k = 0; k < w->max_level && i; ++k) {
        uint64_t o0 = ti_rank(w, offset), oi = ti_rank(w, offset + i) - o0, oe = ti_rank(w, offset + node_size) - o0;
        if (c & mask) {
            offset += node_size - oe;
            node_size = oe;
            i = oi;
        } else {
            node_size -= oe;
            i -= oi;
        }
        offset += w->size;
        mask >>= 1;
    }
    return i;
}

/* wt_int.hpp:418-445 (and operator[] :340-367) */
uint64_t orc_wt_int_inverse_select(const orc_wt_int *w, uint64_t i, uint64_t *sym)
{
    uint64_t c = 0, node_size = w->size, offset = 0;
    uint32_t k;
    for (k = 0; k < w->max_level; ++k) {
        uint64_t o0 = ti_rank(w, offset), oi = ti_rank(w, offset + i) - o0, oe = ti_rank(w, offset + node_size) - o0;
        uint64_t pos = offset + i;
        c <<= 1;
        if ((w->tree[pos >> 6] >> (pos & 63)) & 1) {
            offset += node_size - oe;
            node_size = oe;
            i = oi;
            c |= 1;
        } else {
            node_size -= oe;
            i -= oi;
        }
        offset += w->size;
    }
    *sym = c;
    return i;
}

/* wt_int.hpp:456-507; returns size when c does not occur i times (the reference throws there) */
uint64_t orc_wt_int_select(const orc_wt_int *w, uint64_t i, uint64_t c)
{
    uint64_t offset = 0, mask, node_size = w->size, path_off[65], path_rank_off[65];
    uint32_t k;
    if (w->size == 0 || (w->max_level < 64 && (1ULL << w->max_level) <= c))
        return w->size;
    mask = 1ULL << (w->max_level - 1);
    path_off[0] = path_rank_off[0] = 0;
    for (k = 0; k < w->max_level && node_size; ++k) {
        uint64_t o0 = ti_rank(w, offset), oe = ti_rank(w, offset + node_size) - o0;
        path_rank_off[k] = o0;
        if (c & mask) {
            offset += node_size - oe;
            node_size = oe;
        } else
            node_size -= oe;
        offset += w->size;
        path_off[k + 1] = offset;
        mask >>= 1;
    }
    if (node_size == 0 || node_size < i)
        return w->size;
    mask = 1;
    for (k = w->max_level; k > 0; --k) {
        uint64_t o0 = path_rank_off[k - 1];
        offset = path_off[k - 1];
        if (c & mask)
            i = orc_select_mcl(w->sel1, w->tree, o0 + i) - offset + 1;
        else
            i = orc_select_mcl(w->sel0, w->tree, offset - o0 + i) - offset + 1;
        mask <<= 1;
    }
    return i - 1;
}

void orc_wt_int_rank_batch(const orc_wt_int *w, const uint64_t *i, const uint64_t *c, uint64_t n, uint64_t *out)
{
    uint64_t k;
    for (k = 0; k < n; ++k)
        out[k] = orc_wt_int_rank(w, i[k], c[k]);
}
void orc_wt_int_select_batch(const orc_wt_int *w, const uint64_t *i, const uint64_t *c, uint64_t n, uint64_t *out)
{
    uint64_t k;
    for (k = 0; k < n; ++k)
        out[k] = orc_wt_int_select(w, i[k], c[k]);
}
void orc_wt_int_access_batch(const orc_wt_int *w, const uint64_t *i, uint64_t n, uint64_t *sym, uint64_t *rnk)
{
    uint64_t k, r;
    for (k = 0; k < n; ++k) {
        r = orc_wt_int_inverse_select(w, i[k], &sym[k]);
        if (rnk)
            rnk[k] = r;
    }
}

/* wt_int.hpp:792-805 */
uint64_t orc_wt_int_serialize(const orc_wt_int *w, uint8_t *out, uint64_t cap)
{
    orc_buf b = {0, 0, 0};
    orc__buf_u64(&b, w->size);
    orc__buf_u64(&b, w->sigma);
    orc__bv_serialize_into(&b, w->tree, w->tree_bits);
    orc__rank_v_serialize_into(&b, w->rank_table, w->tree_bits);
    orc__select_mcl_serialize_into(&b, w->sel1);
    orc__select_mcl_serialize_into(&b, w->sel0);
    orc__buf_put(&b, &w->max_level, 4);
    return orc__buf_finish(&b, out, cap);
}
