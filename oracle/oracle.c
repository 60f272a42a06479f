/* oracle/oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See oracle.h for the contract.
 *
 * Part 1: bits (a1), int_vector packing/serialisation (a2), rank_support_v (a3),
 *         select_support_mcl (a4).
 * Citations are relative to /root/reference/include/sdsl/.
 */
#include "oracle_priv.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* a1: bits.hpp                                                                               */
/* ------------------------------------------------------------------------------------------ */

/* bits.hpp:486-502 — population count (SWAR form of the non-SSE path, :492-500) */
uint32_t orc_cnt(uint64_t x)
{
    x = x - ((x >> 1) & 0x5555555555555555ULL);
    x = (x & 0x3333333333333333ULL) + ((x >> 2) & 0x3333333333333333ULL);
    x = (x + (x >> 4)) & 0x0f0f0f0f0f0f0f0fULL;
    return (uint32_t)((x * 0x0101010101010101ULL) >> 56);
}

/* bits.hpp:586-612 — position of the i-th (1-based) set bit; restated as "drop i-1 lowest ones" */
uint32_t orc_sel(uint64_t x, uint32_t i)
{
    while (--i)
        x &= x - 1;
    return orc_lo(x);
}

/* bits.hpp:653-684 — index of the most significant set bit, hi(0) = 0 */
uint32_t orc_hi(uint64_t x)
{
    uint32_t r = 0;
    while (x >>= 1)
        ++r;
    return r;
}

/* bits.hpp:689-709 — index of the least significant set bit, lo(0) = 0 */
uint32_t orc_lo(uint64_t x)
{
    uint32_t r = 0;
    if (x == 0)
        return 0;
    while (!(x & 1)) {
        x >>= 1;
        ++r;
    }
    return r;
}

uint64_t orc__lo_set(uint32_t k) /* bits.hpp:194-211 */
{
    return k >= 64 ? ~0ULL : ((1ULL << k) - 1);
}

/* bits.hpp:777-790 — read `len` (<= 64) bits starting at absolute bit position `bitpos` */
uint64_t orc_read_int(const uint64_t *d, uint64_t bitpos, uint8_t len)
{
    const uint64_t *w = d + (bitpos >> 6);
    uint32_t off = (uint32_t)(bitpos & 63);
    if (len == 0)
        return 0;
    if (off + len > 64) {
        uint64_t lo = w[0] >> off;
        uint64_t hi = w[1] & orc__lo_set(off + len - 64);
        return lo | (hi << (64 - off));
    }
    return (w[0] >> off) & orc__lo_set(len);
}

void orc__write_int(uint64_t *d, uint64_t bitpos, uint64_t x, uint8_t len) /* bits.hpp:748-773 */
{
    uint64_t *w = d + (bitpos >> 6);
    uint32_t off = (uint32_t)(bitpos & 63);
    if (len == 0)
        return;
    x &= orc__lo_set(len);
    if (off + len > 64) {
        w[0] = (w[0] & orc__lo_set(off)) | (x << off);
        w[1] = (w[1] & ~orc__lo_set(off + len - 64)) | (x >> (64 - off));
    } else {
        uint64_t m = orc__lo_set(len) << off;
        w[0] = (w[0] & ~m) | (x << off);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* byte buffer + int_vector                                                                   */
/* ------------------------------------------------------------------------------------------ */

void orc__buf_put(orc_buf *b, const void *src, uint64_t n)
{
    if (b->n + n > b->cap) {
        uint64_t c = b->cap ? b->cap * 2 : 4096;
        while (c < b->n + n)
            c *= 2;
        b->p = (uint8_t *)realloc(b->p, c);
        b->cap = c;
    }
    memcpy(b->p + b->n, src, n);
    b->n += n;
}
void orc__buf_u64(orc_buf *b, uint64_t x)
{
    orc__buf_put(b, &x, 8);
}
void orc_buf_free(orc_buf *b)
{
    free(b->p);
    b->p = NULL;
    b->n = b->cap = 0;
}
uint64_t orc__buf_finish(orc_buf *b, uint8_t *out, uint64_t cap)
{
    uint64_t n = b->n;
    if (out != NULL && n <= cap)
        memcpy(out, b->p, n);
    orc_buf_free(b);
    return n;
}

void orc__iv_init(orc_iv *v, uint64_t size, uint8_t width)
{
    uint64_t words = ((size * width + 63) >> 6) + 1;
    v->size = size;
    v->width = width;
    v->data = (uint64_t *)calloc(words, 8);
}
void orc_iv_free(orc_iv *v)
{
    free(v->data);
    v->data = NULL;
    v->size = 0;
}
uint64_t orc_iv_get(const orc_iv *v, uint64_t i) /* int_vector.hpp:1865-1869 */
{
    return orc_read_int(v->data, i * v->width, v->width);
}
void orc__iv_set(orc_iv *v, uint64_t i, uint64_t x)
{
    orc__write_int(v->data, i * v->width, x, v->width);
}
/* int_vector.hpp:904-916,1995-2004: header (width<<56 | bit_size) then ceil(bit_size/64) words */
void orc__iv_serialize(orc_buf *b, const orc_iv *v)
{
    uint64_t bits = v->size * v->width;
    orc__buf_u64(b, ((uint64_t)v->width << 56) | bits);
    if (bits)
        orc__buf_put(b, v->data, ((bits + 63) >> 6) * 8);
}
static void empty_iv_serialize(orc_buf *b) /* default int_vector<0>: width 64, size 0 */
{
    orc__buf_u64(b, (uint64_t)64 << 56);
}

void orc__bv_serialize_into(orc_buf *b, const uint64_t *w, uint64_t nbits)
{
    orc__buf_u64(b, ((uint64_t)1 << 56) | nbits);
    if (nbits)
        orc__buf_put(b, w, ((nbits + 63) >> 6) * 8);
}
uint64_t orc_bv_serialize(const uint64_t *w, uint64_t nbits, uint8_t *out, uint64_t cap)
{
    orc_buf b = {0, 0, 0};
    orc__bv_serialize_into(&b, w, nbits);
    return orc__buf_finish(&b, out, cap);
}

/* ------------------------------------------------------------------------------------------ */
/* a3: rank_support_v<b,1>                                                                    */
/* ------------------------------------------------------------------------------------------ */

/* The word of `pattern ends here` indicator bits every trait function of the reference is a popcount / select of:
 *   pattern codes b: 0, 1 (one bit, rank_support.hpp:105-158, select_support.hpp:120-201) and the two-bit patterns
 *   2 = "10", 3 = "01", 4 = "00", 5 = "11" (rank_support.hpp:161-284, select_support.hpp:204-405; bits.hpp:565-583).
 * carry = msb of the previous word, or init_carry() for word 0 (rank_support.hpp:184-187,214-217,247-250,280-283;
 * select_support.hpp:239-242,276-279,339-342,398-401): "01" and "00" start with 1, so bit 0 never ends a pattern. */
static uint64_t pat_map(const uint64_t *w, uint64_t k, int b)
{
    uint64_t x = w[k], c;
    if (b < 2)
        return b ? x : ~x;
    c = k ? (w[k - 1] >> 63) : (uint64_t)(b == 3 || b == 4);
    switch (b) {
    case 2:
        return ((x << 1) | c) & ~x; /* map10 */
    case 3:
        return (x ^ ((x << 1) | c)) & x; /* map01 */
    case 4:
        return ~(x | ((x << 1) | c));
    default:
        return x & ((x << 1) | c);
    }
}
static uint32_t args(const uint64_t *w, uint64_t k, int b) /* args_in_the_word */
{
    return orc_cnt(pat_map(w, k, b));
}

uint64_t orc_rank_v_table_words(uint64_t nbits) /* rank_support_v.hpp:79,84 */
{
    if (nbits == 0)
        return 2;
    return (((nbits + 63) >> 9) + 1) << 1;
}

/* rank_support_v.hpp:72-122 */
void orc_rank_v_build(const uint64_t *w, uint64_t nbits, int b, uint64_t *B)
{
    uint64_t W = (nbits + 63) >> 6, i, j = 0, sum, second = 0;
    B[0] = B[1] = 0;
    if (nbits == 0)
        return;
    sum = args(w, 0, b);
    for (i = 1; i < W; ++i) {
        if ((i & 7) == 0) {
            j += 2;
            B[j - 1] = second;
            B[j] = B[j - 2] + sum;
            second = sum = 0;
        } else {
            second |= sum << (63 - 9 * (i & 7));
        }
        sum += args(w, i, b);
    }
    if (i & 7) {
        second |= sum << (63 - 9 * (i & 7));
        B[j + 1] = second;
    } else {
        j += 2;
        B[j - 1] = second;
        B[j] = B[j - 2] + sum;
        B[j + 1] = 0;
    }
}

/* rank_support_v.hpp:129-139 with word_rank of rank_support.hpp:120-123,146-149 */
uint64_t orc_rank_v(const uint64_t *w, const uint64_t *B, int b, uint64_t idx)
{
    const uint64_t *p = B + ((idx >> 8) & ~1ULL);
    uint64_t r = p[0] + ((p[1] >> (63 - 9 * ((idx & 0x1FF) >> 6))) & 0x1FF);
    if (idx & 0x3F)
        r += orc_cnt(pat_map(w, idx >> 6, b) & orc__lo_set((uint32_t)(idx & 0x3F)));
    return r;
}

void orc_rank_v_batch(const uint64_t *w, const uint64_t *B, int b, const uint64_t *idx, uint64_t n, uint64_t *out)
{
    uint64_t k;
    for (k = 0; k < n; ++k)
        out[k] = orc_rank_v(w, B, b, idx[k]);
}

void orc__rank_v_serialize_into(orc_buf *b, const uint64_t *B, uint64_t nbits)
{
    uint64_t words = orc_rank_v_table_words(nbits);
    orc__buf_u64(b, ((uint64_t)64 << 56) | (words * 64));
    orc__buf_put(b, B, words * 8);
}
uint64_t orc_rank_v_serialize(const uint64_t *B, uint64_t nbits, uint8_t *out, uint64_t cap)
{
    orc_buf b = {0, 0, 0};
    orc__rank_v_serialize_into(&b, B, nbits);
    return orc__buf_finish(&b, out, cap);
}

/* ------------------------------------------------------------------------------------------ */
/* f-3: rank_support_v5<b,1> — the 6.25 %-overhead table the reference's own count benchmark    */
/* uses (benchmark/indexing_count/index.config:8): 2048-bit superblocks, per superblock the       */
/* absolute count and five 12-bit prefix counts, one per 6 words                                  */
/* ------------------------------------------------------------------------------------------ */
uint64_t orc_rank_v5_table_words(uint64_t nbits) /* rank_support_v5.hpp:73-79 */
{
    if (nbits == 0)
        return 2;
    return (((nbits + 63) >> 11) + 1) << 1;
}

/* rank_support_v5.hpp:66-122 */
void orc_rank_v5_build(const uint64_t *w, uint64_t nbits, int b, uint64_t *B)
{
    uint64_t W = (nbits + 63) >> 6, i, j = 0, sum, second = 0, cw = 1;
    B[0] = B[1] = 0;
    if (nbits == 0)
        return;
    sum = args(w, 0, b);
    for (i = 1; i < W; ++i, ++cw) {
        if (cw == 32) {
            j += 2;
            B[j - 1] = second;
            B[j] = B[j - 2] + sum;
            second = sum = cw = 0;
        } else if (cw % 6 == 0) {
            second |= sum << (60 - 12 * (cw / 6));
        }
        sum += args(w, i, b);
    }
    if (cw % 6 == 0)
        second |= sum << (60 - 12 * (cw / 6));
    if (cw == 32) {
        j += 2;
        B[j - 1] = second;
        B[j] = B[j - 2] + sum;
        B[j + 1] = 0;
    } else {
        B[j + 1] = second;
    }
}

/* rank_support_v5.hpp:131-149 */
uint64_t orc_rank_v5(const uint64_t *w, const uint64_t *B, int b, uint64_t idx)
{
    const uint64_t *p = B + ((idx >> 10) & ~1ull);
    uint64_t r = p[0] + ((p[1] >> (60 - 12 * ((idx & 0x7FF) / 384))) & 0x7FF);
    uint64_t word = idx >> 6, todo = (word & 31) % 6, k;
    if (idx & 63)
        r += orc_cnt(pat_map(w, word, b) & orc__lo_set((uint32_t)(idx & 63)));
    for (k = 1; k <= todo; ++k)
        r += args(w, word - k, b);
    return r;
}

uint64_t orc_rank_v5_serialize(const uint64_t *B, uint64_t nbits, uint8_t *out, uint64_t cap) /* :151-158 */
{
    orc_buf b = {0, 0, 0};
    uint64_t words = orc_rank_v5_table_words(nbits);
    orc__buf_u64(&b, ((uint64_t)64 << 56) | (words * 64));
    orc__buf_put(&b, B, words * 8);
    return orc__buf_finish(&b, out, cap);
}

/* ------------------------------------------------------------------------------------------ */
/* a4: select_support_mcl<b,1>                                                                */
/* ------------------------------------------------------------------------------------------ */

static int found_arg(const uint64_t *w, uint64_t i, int b) /* select_support.hpp:149-152,191-194,233-238,270-275,334-337,392-397 */
{
    int cur = (int)((w[i >> 6] >> (i & 63)) & 1), prev;
    if (b < 2)
        return cur == b;
    if (i == 0)
        return 0;
    prev = (int)((w[(i - 1) >> 6] >> ((i - 1) & 63)) & 1);
    return prev == (b == 2 || b == 5) && cur == (b == 3 || b == 5);
}

#define SBS 4096u

/* select_support_mcl.hpp:207-266 */
static void init_slow(orc_selmcl *s, const uint64_t *w)
{
    uint64_t *pos = (uint64_t *)malloc(SBS * 8);
    uint64_t cnt = 0, sbc = 0, i;
    for (i = 0; i < s->nbits; ++i) {
        if (!found_arg(w, i, s->b))
            continue;
        pos[cnt % SBS] = i;
        ++cnt;
        if (cnt % SBS == 0 || cnt == s->arg_cnt) {
            uint64_t last = (cnt - 1) % SBS, j;
            uint64_t diff = pos[last] - pos[0];
            orc__iv_set(&s->superblock, sbc, pos[0]);
            if (diff > s->logn4) {
                s->has_long = 1;
                orc__iv_init(&s->longsb[sbc], SBS, (uint8_t)(orc_hi(pos[last]) + 1));
                for (j = 0; j <= last; ++j)
                    orc__iv_set(&s->longsb[sbc], j, pos[j]);
            } else {
                orc__iv_init(&s->mini[sbc], 64, (uint8_t)(orc_hi(diff) + 1));
                for (j = 0; j <= last; j += 64)
                    orc__iv_set(&s->mini[sbc], j / 64, pos[j] - pos[0]);
            }
            ++sbc;
        }
    }
    free(pos);
}

/* select_support_mcl.hpp:269-381 — word-wise builder; note its quirks, all preserved:
 *  - the "last arg of the block" it measures is really the FIRST arg of the next superblock (:311-318);
 *  - the trailing partial superblock is always stored long with width hi(n-1)+1 and its
 *    m_superblock entry stays 0 (:366-380);
 *  - for b == 0 the running count is clamped to arg_cnt so padding zeros are never sampled (:299-300). */
static void init_fast(orc_selmcl *s, const uint64_t *w)
{
    uint64_t *pos = (uint64_t *)calloc(SBS, 8);
    uint64_t last_k64 = 1, sbc = 0, i, cnt_old = 0, cnt_new = 0, last_k64_sum = 1;
    uint64_t nb64 = ((s->nbits + 63) >> 6) << 6;
    for (i = 0; i < nb64; i += 64) {
        uint64_t x = s->b ? w[i >> 6] : ~w[i >> 6];
        cnt_new += orc_cnt(x);
        if (cnt_new > s->arg_cnt)
            cnt_new = s->arg_cnt;
        if (cnt_new >= last_k64_sum) {
            pos[last_k64 - 1] = i + orc_sel(x, (uint32_t)(last_k64_sum - cnt_old));
            last_k64 += 64;
            last_k64_sum += 64;
            if (last_k64 == SBS + 1) {
                uint64_t plast = pos[last_k64 - 65], ii, j, k, diff;
                orc__iv_set(&s->superblock, sbc, pos[0]);
                for (ii = pos[last_k64 - 65] + 1, j = last_k64 - 65; ii < s->nbits && j < SBS; ++ii)
                    if (found_arg(w, ii, s->b)) {
                        plast = ii;
                        ++j;
                    }
                diff = plast - pos[0];
                if (diff > s->logn4) {
                    s->has_long = 1;
                    orc__iv_init(&s->longsb[sbc], SBS, (uint8_t)(orc_hi(plast) + 1));
                    for (j = pos[0], k = 0; k < SBS && j <= plast; ++j)
                        if (found_arg(w, j, s->b))
                            orc__iv_set(&s->longsb[sbc], k++, j);
                } else {
                    orc__iv_init(&s->mini[sbc], 64, (uint8_t)(orc_hi(diff) + 1));
                    for (j = 0; j < SBS; j += 64)
                        orc__iv_set(&s->mini[sbc], j / 64, pos[j] - pos[0]);
                }
                ++sbc;
                last_k64 = 1;
            }
        }
        cnt_old = cnt_new;
    }
    if (last_k64 > 1) {
        uint64_t k = 0;
        s->has_long = 1;
        orc__iv_init(&s->longsb[sbc], SBS, (uint8_t)(orc_hi(s->nbits - 1) + 1));
        for (i = pos[0]; i < s->nbits; ++i)
            if (found_arg(w, i, s->b))
                orc__iv_set(&s->longsb[sbc], k++, i);
        ++sbc;
    }
    free(pos);
}

orc_selmcl *orc_select_mcl_build(const uint64_t *w, uint64_t nbits, int b)
{
    orc_selmcl *s = (orc_selmcl *)calloc(1, sizeof(*s));
    uint64_t W = (nbits + 63) >> 6, k, ones = 0;
    s->nbits = nbits;
    s->b = b;
    /* initData, select_support_mcl.hpp:448-465 (uint32 arithmetic) */
    s->logn = orc_hi(W << 6) + 1;
    s->logn2 = s->logn * s->logn;
    s->logn4 = s->logn2 * s->logn2;
    /* arg_cnt: select_support.hpp:127-130,169-172 (cnt_one_bits ignores bits past n, util.hpp) */
    for (k = 0; k < W; ++k) {
        uint64_t x = w[k];
        if (k == W - 1 && (nbits & 63))
            x &= orc__lo_set((uint32_t)(nbits & 63));
        ones += orc_cnt(x);
    }
    s->arg_cnt = b ? ones : nbits - ones;
    if (b >= 2) /* cnt_onezero_bits / cnt_zeroone_bits (util.hpp:689-726), arg_cnt of "00"/"11" (select_support.hpp:286-301,350-365) */
        for (s->arg_cnt = 0, k = 0; k < nbits; ++k)
            s->arg_cnt += (uint64_t)found_arg(w, k, b);
    if (s->arg_cnt == 0)
        return s;
    s->sb = (s->arg_cnt + SBS - 1) / SBS;
    orc__iv_init(&s->superblock, s->sb, (uint8_t)s->logn);
    s->longsb = (orc_iv *)calloc(s->sb + 1, sizeof(orc_iv));
    s->mini = (orc_iv *)calloc(s->sb + 1, sizeof(orc_iv));
    if (b >= 2 || nbits < 100000) /* ctor dispatch, :121-128: two-bit patterns always take init_slow */
        init_slow(s, w);
    else
        init_fast(s, w);
    return s;
}

void orc_select_mcl_free(orc_selmcl *s)
{
    uint64_t k;
    if (!s)
        return;
    for (k = 0; k < s->sb; ++k) {
        if (s->longsb)
            orc_iv_free(&s->longsb[k]);
        if (s->mini)
            orc_iv_free(&s->mini[k]);
    }
    free(s->longsb);
    free(s->mini);
    orc_iv_free(&s->superblock);
    free(s);
}

/* select_support_mcl.hpp:384-439 */
uint64_t orc_select_mcl(const orc_selmcl *s, const uint64_t *w, uint64_t i)
{
    uint64_t sb_idx, offset, pos, wp, sum, x;
    uint32_t wo, a;
    i -= 1;
    sb_idx = i >> 12;
    offset = i & 0xFFF;
    if (s->has_long && s->longsb[sb_idx].size != 0)
        return orc_iv_get(&s->longsb[sb_idx], offset);
    pos = orc_iv_get(&s->superblock, sb_idx) + orc_iv_get(&s->mini[sb_idx], offset >> 6);
    if ((offset & 0x3F) == 0)
        return pos;
    i = offset & 0x3F; /* args still to find, 1..63 */
    pos += 1;
    wp = pos >> 6;
    wo = (uint32_t)(pos & 63);
    x = pat_map(w, wp, s->b) & ~orc__lo_set(wo); /* args_in_the_first_word with init_carry(data, word_pos) */
    a = orc_cnt(x);
    if (a >= i)
        return (wp << 6) + orc_sel(x, (uint32_t)i);
    sum = a;
    for (;;) {
        ++wp;
        x = pat_map(w, wp, s->b); /* args_in_the_word with get_carry of the previous word */
        a = orc_cnt(x);
        if (sum + a >= i)
            return (wp << 6) + orc_sel(x, (uint32_t)(i - sum));
        sum += a;
    }
}

void orc_select_mcl_batch(const orc_selmcl *s, const uint64_t *w, const uint64_t *i, uint64_t n, uint64_t *out)
{
    uint64_t k;
    for (k = 0; k < n; ++k)
        out[k] = orc_select_mcl(s, w, i[k]);
}

/* select_support_mcl.hpp:474-518 */
void orc__select_mcl_serialize_into(orc_buf *bp, const orc_selmcl *s)
{
    uint64_t k;
    orc__buf_u64(bp, s->arg_cnt);
    if (s->arg_cnt) {
        orc_iv mol;
        orc__iv_serialize(bp, &s->superblock);
        if (s->has_long) {
            orc__iv_init(&mol, s->sb, 1);
            for (k = 0; k < s->sb; ++k)
                orc__iv_set(&mol, k, s->mini[k].size != 0);
            orc__iv_serialize(bp, &mol);
        } else {
            orc__iv_init(&mol, 0, 1);
            orc__iv_serialize(bp, &mol);
        }
        for (k = 0; k < s->sb; ++k) {
            if (s->has_long && !orc_iv_get(&mol, k)) {
                if (s->longsb[k].size)
                    orc__iv_serialize(bp, &s->longsb[k]);
                else
                    empty_iv_serialize(bp);
            } else {
                if (s->mini[k].size)
                    orc__iv_serialize(bp, &s->mini[k]);
                else
                    empty_iv_serialize(bp);
            }
        }
        orc_iv_free(&mol);
    }
}
uint64_t orc_select_mcl_serialize(const orc_selmcl *s, uint8_t *out, uint64_t cap)
{
    orc_buf b = {0, 0, 0};
    orc__select_mcl_serialize_into(&b, s);
    return orc__buf_finish(&b, out, cap);
}
