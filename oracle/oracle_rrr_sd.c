/* oracle/oracle_rrr_sd.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See oracle.h.
 *
 * Part 4: rrr_vector<63, int_vector<>, 32> (row a5) and sd_vector<> (row a6) with their rank / select
 * supports: construction, queries and the reference's serialised form.
 * Citations are relative to /root/reference/include/sdsl/.
 */
#include "oracle_priv.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* a5: rrr_vector<63>                                                                          */
/* ------------------------------------------------------------------------------------------ */
#define BS 63u /* t_bs */
#define KK 32u /* t_k  */

static uint64_t binom[65][65]; /* rrr_helper.hpp:193-237 */
static uint8_t space[64];      /* rrr_helper.hpp:286-293: bits to store an offset of class k */
static int tables_ready = 0;

static void init_tables(void)
{
    int nn, k;
    if (tables_ready)
        return;
    memset(binom, 0, sizeof(binom));
    for (nn = 0; nn <= 64; ++nn)
        binom[nn][0] = 1;
    for (nn = 1; nn <= 64; ++nn)
        for (k = 1; k <= nn; ++k)
            binom[nn][k] = binom[nn - 1][k - 1] + binom[nn - 1][k];
    for (k = 0; k <= 63; ++k)
        space[k] = (binom[63][k] == 1) ? 0 : (uint8_t)(orc_hi(binom[63][k]) + 1);
    tables_ready = 1;
}

/* rrr_helper.hpp:346-366 */
static uint64_t bin_to_nr(uint64_t bin)
{
    uint64_t nr = 0;
    uint32_t k = orc_cnt(bin), nn = BS;
    if (bin == 0 || bin == orc__lo_set(BS))
        return 0;
    while (bin) {
        if (bin & 1) {
            nr += binom[nn - 1][k];
            --k;
        }
        bin >>= 1;
        --nn;
    }
    return nr;
}

/* inverse of bin_to_nr: the 63-bit block with k ones and offset nr.  Every decode_* routine of
 * rrr_helper.hpp (:369-649: decode_bit, decode_popcount, decode_select, decode_select_bitpattern) returns a
 * function of this block; restating the block once keeps the oracle independent of their control flow. */
static uint64_t nr_to_bin(uint32_t k, uint64_t nr)
{
    uint64_t bin = 0;
    uint32_t nn, p = 0;
    if (k == 0)
        return 0;
    if (k == BS)
        return orc__lo_set(BS);
    for (nn = BS; nn > 0 && k > 0; --nn, ++p)
        if (nr >= binom[nn - 1][k]) {
            nr -= binom[nn - 1][k];
            bin |= 1ULL << p;
            --k;
        }
    return bin;
}

/* rrr_vector.hpp:158-270 */
orc_rrr *orc_rrr_build(const uint64_t *w, uint64_t nbits)
{
    orc_rrr *r = (orc_rrr *)calloc(1, sizeof(*r));
    uint64_t nb = (nbits + BS) / BS, pos, i, btnr_pos = 0, sum_rank = 0, nsb;
    init_tables();
    r->size = nbits;
    orc__iv_init(&r->bt, nb, 6); /* width hi(63)+1 */
    for (pos = 0, i = 0; pos + BS <= nbits; pos += BS) {
        uint32_t x = orc_cnt(orc_read_int(w, pos, BS));
        orc__iv_set(&r->bt, i++, x);
        sum_rank += x;
        btnr_pos += space[x];
    }
    if (pos < nbits) {
        uint32_t x = orc_cnt(orc_read_int(w, pos, (uint8_t)(nbits - pos)));
        orc__iv_set(&r->bt, i++, x);
        sum_rank += x;
        btnr_pos += space[x];
    }
    nsb = (nb + KK - 1) / KK;
    r->btnr_bits = btnr_pos > 64 ? btnr_pos : 64;
    r->btnr = (uint64_t *)calloc(((r->btnr_bits + 63) >> 6) + 2, 8);
    orc__iv_init(&r->btnrp, nsb, (uint8_t)(orc_hi(btnr_pos) + 1));
    orc__iv_init(&r->rank, nsb + ((nbits % (KK * BS)) > 0), (uint8_t)(orc_hi(sum_rank) + 1));
    orc__iv_init(&r->invert, nsb, 1);
    {
        int inv = 0;
        btnr_pos = 0;
        sum_rank = 0;
        for (pos = 0, i = 0; pos < nbits; pos += BS) {
            uint32_t len = (pos + BS <= nbits) ? BS : (uint32_t)(nbits - pos), x, sp;
            if (i % KK == 0) {
                orc__iv_set(&r->btnrp, i / KK, btnr_pos);
                orc__iv_set(&r->rank, i / KK, sum_rank);
                inv = 0;
                if (len == BS && i + KK <= nb) { /* invert bit only for complete superblocks (:203-228) */
                    uint64_t j, gt = 0;
                    for (j = i; j < i + KK; ++j)
                        gt += orc_iv_get(&r->bt, j) > BS / 2;
                    if (gt > KK / 2) {
                        orc__iv_set(&r->invert, i / KK, 1);
                        for (j = i; j < i + KK; ++j)
                            orc__iv_set(&r->bt, j, BS - orc_iv_get(&r->bt, j));
                        inv = 1;
                    }
                }
            }
            x = (uint32_t)orc_iv_get(&r->bt, i++);
            sp = space[x];
            sum_rank += inv ? BS - x : x;
            if (sp)
                orc__write_int(r->btnr, btnr_pos, bin_to_nr(orc_read_int(w, pos, (uint8_t)len)), (uint8_t)sp);
            btnr_pos += sp;
        }
    }
    orc__iv_set(&r->rank, r->rank.size - 1, sum_rank);
    return r;
}

void orc_rrr_free(orc_rrr *r)
{
    if (!r)
        return;
    orc_iv_free(&r->bt);
    orc_iv_free(&r->btnrp);
    orc_iv_free(&r->rank);
    orc_iv_free(&r->invert);
    free(r->btnr);
    free(r);
}

/* the (de-inverted) class and the decoded bits of block b */
static uint64_t rrr_block(const orc_rrr *r, uint64_t b, uint32_t *k_out)
{
    uint64_t g = b / KK, p = orc_iv_get(&r->btnrp, g), j;
    int inv = (int)orc_iv_get(&r->invert, g);
    uint32_t k;
    for (j = g * KK; j < b; ++j)
        p += space[orc_iv_get(&r->bt, j)];
    k = (uint32_t)orc_iv_get(&r->bt, b);
    if (inv)
        k = BS - k;
    *k_out = k;
    return nr_to_bin(k, orc_read_int(r->btnr, p, space[k]));
}

/* rrr_vector.hpp:503-544 */
uint64_t orc_rrr_rank(const orc_rrr *r, int b, uint64_t i)
{
    uint64_t blk = i / BS, g = blk / KK, rank = orc_iv_get(&r->rank, g), j;
    uint32_t off = (uint32_t)(i % BS), k;
    int inv = (int)orc_iv_get(&r->invert, g);
    for (j = g * KK; j < blk; ++j) {
        uint64_t c = orc_iv_get(&r->bt, j);
        rank += inv ? BS - c : c;
    }
    if (off)
        rank += orc_cnt(rrr_block(r, blk, &k) & orc__lo_set(off));
    return b ? rank : i - rank;
}

/* rrr_vector.hpp:276-298 */
uint64_t orc_rrr_access(const orc_rrr *r, uint64_t i)
{
    uint32_t k;
    return (rrr_block(r, i / BS, &k) >> (i % BS)) & 1;
}

/* rrr_vector.hpp:639-726: i beyond the number of b-bits returns size() */
uint64_t orc_rrr_select(const orc_rrr *r, int b, uint64_t i)
{
    uint64_t total1 = orc_iv_get(&r->rank, r->rank.size - 1), begin = 0, end = r->rank.size - 1, cnt, idx;
    uint32_t k = 0;
    if ((b ? total1 : r->size - total1) < i)
        return r->size;
    while (end - begin > 1) { /* superblock g with count_before(g) < i <= count_before(g+1) */
        uint64_t mid = (begin + end) >> 1;
        uint64_t c = b ? orc_iv_get(&r->rank, mid) : mid * BS * KK - orc_iv_get(&r->rank, mid);
        if (c >= i)
            end = mid;
        else
            begin = mid;
    }
    cnt = b ? orc_iv_get(&r->rank, begin) : begin * BS * KK - orc_iv_get(&r->rank, begin);
    for (idx = begin * KK;; ++idx) {
        uint64_t bits = rrr_block(r, idx, &k);
        uint64_t x = b ? bits : (~bits & orc__lo_set(BS));
        uint32_t c = orc_cnt(x);
        if (cnt + c >= i)
            return idx * BS + orc_sel(x, (uint32_t)(i - cnt));
        cnt += c;
    }
}

void orc_rrr_rank_batch(const orc_rrr *r, int b, const uint64_t *i, uint64_t n, uint64_t *out)
{
    uint64_t k;
    for (k = 0; k < n; ++k)
        out[k] = orc_rrr_rank(r, b, i[k]);
}
void orc_rrr_select_batch(const orc_rrr *r, int b, const uint64_t *i, uint64_t n, uint64_t *out)
{
    uint64_t k;
    for (k = 0; k < n; ++k)
        out[k] = orc_rrr_select(r, b, i[k]);
}
void orc_rrr_access_batch(const orc_rrr *r, const uint64_t *i, uint64_t n, uint64_t *out)
{
    uint64_t k;
    for (k = 0; k < n; ++k)
        out[k] = orc_rrr_access(r, i[k]);
}

/* rrr_vector.hpp:366-378 */
uint64_t orc_rrr_serialize(const orc_rrr *r, uint8_t *out, uint64_t cap)
{
    orc_buf b = {0, 0, 0};
    orc__buf_u64(&b, r->size);
    orc__iv_serialize(&b, &r->bt);
    orc__bv_serialize_into(&b, r->btnr, r->btnr_bits);
    orc__iv_serialize(&b, &r->btnrp);
    orc__iv_serialize(&b, &r->rank);
    orc__iv_serialize(&b, &r->invert);
    return orc__buf_finish(&b, out, cap);
}

/* ------------------------------------------------------------------------------------------ */
/* a6: sd_vector<>                                                                             */
/* ------------------------------------------------------------------------------------------ */

/* sd_vector.hpp:218-257 */
orc_sd *orc_sd_build(const uint64_t *w, uint64_t nbits)
{
    orc_sd *s = (orc_sd *)calloc(1, sizeof(*s));
    uint64_t W = (nbits + 63) >> 6, k, m = 0, mm = 0, last_high = 0, highpos = 0, i;
    uint8_t logm, logn;
    for (k = 0; k < W; ++k) {
        uint64_t x = w[k];
        if (k == W - 1 && (nbits & 63))
            x &= orc__lo_set((uint32_t)(nbits & 63));
        m += orc_cnt(x);
    }
    s->size = nbits;
    s->m = m;
    logm = (uint8_t)(orc_hi(m) + 1);
    logn = (uint8_t)(orc_hi(nbits) + 1);
    if (logm == logn)
        --logm;
    s->wl = (uint8_t)(logn - logm);
    orc__iv_init(&s->low, m, s->wl);
    s->high_bits = m + (1ULL << logm);
    s->high = (uint64_t *)calloc(((s->high_bits + 63) >> 6) + 2, 8);
    for (i = 0; i < nbits; ++i) {
        if (!((w[i >> 6] >> (i & 63)) & 1))
            continue;
        {
            uint64_t cur_high = i >> s->wl;
            highpos += cur_high - last_high;
            last_high = cur_high;
            orc__iv_set(&s->low, mm++, i); /* truncated to wl bits */
            s->high[highpos >> 6] |= 1ULL << (highpos & 63);
            ++highpos;
        }
    }
    s->sel1 = orc_select_mcl_build(s->high, s->high_bits, 1);
    s->sel0 = orc_select_mcl_build(s->high, s->high_bits, 0);
    return s;
}

void orc_sd_free(orc_sd *s)
{
    if (!s)
        return;
    orc_iv_free(&s->low);
    free(s->high);
    orc_select_mcl_free(s->sel1);
    orc_select_mcl_free(s->sel0);
    free(s);
}

static int high_bit(const orc_sd *s, uint64_t p)
{
    return (int)((s->high[p >> 6] >> (p & 63)) & 1);
}

/* sd_vector.hpp:553-575 */
uint64_t orc_sd_rank(const orc_sd *s, int b, uint64_t i)
{
    uint64_t hv = i >> s->wl, sh = orc_select_mcl(s->sel0, s->high, hv + 1), rl = sh - hv, vl, r;
    if (rl == 0)
        r = 0;
    else {
        vl = i & orc__lo_set(s->wl);
        r = ~0ULL;
        do {
            if (!sh) {
                r = 0;
                break;
            }
            --sh;
            --rl;
        } while (high_bit(s, sh) && orc_iv_get(&s->low, rl) >= vl);
        if (r == ~0ULL)
            r = rl + 1;
    }
    return b ? r : i - r;
}

/* sd_vector.hpp:621-630 */
static uint64_t sd_select1(const orc_sd *s, uint64_t i)
{
    return orc_iv_get(&s->low, i - 1) + ((orc_select_mcl(s->sel1, s->high, i) + 1 - i) << s->wl);
}

/* sd_vector.hpp:632-664 */
uint64_t orc_sd_select(const orc_sd *s, int b, uint64_t i)
{
    uint64_t lb = 1, rb = s->m + 1, r0 = 0, pos = ~0ULL;
    if (b)
        return sd_select1(s, i);
    while (lb < rb) {
        uint64_t mid = lb + (rb - lb) / 2, x = sd_select1(s, mid), rank0 = x + 1 - mid;
        if (rank0 >= i)
            rb = mid;
        else {
            r0 = rank0;
            pos = x;
            lb = mid + 1;
        }
    }
    return pos + i - r0;
}

/* sd_vector.hpp:328-349 */
uint64_t orc_sd_access(const orc_sd *s, uint64_t i)
{
    return orc_sd_rank(s, 1, i + 1) - orc_sd_rank(s, 1, i);
}

void orc_sd_rank_batch(const orc_sd *s, int b, const uint64_t *i, uint64_t n, uint64_t *out)
{
    uint64_t k;
    for (k = 0; k < n; ++k)
        out[k] = orc_sd_rank(s, b, i[k]);
}
void orc_sd_select_batch(const orc_sd *s, int b, const uint64_t *i, uint64_t n, uint64_t *out)
{
    uint64_t k;
    for (k = 0; k < n; ++k)
        out[k] = orc_sd_select(s, b, i[k]);
}
void orc_sd_access_batch(const orc_sd *s, const uint64_t *i, uint64_t n, uint64_t *out)
{
    uint64_t k;
    for (k = 0; k < n; ++k)
        out[k] = orc_sd_access(s, i[k]);
}

/* sd_vector.hpp:426-438 */
uint64_t orc_sd_serialize(const orc_sd *s, uint8_t *out, uint64_t cap)
{
    orc_buf b = {0, 0, 0};
    orc__buf_u64(&b, s->size);
    orc__buf_put(&b, &s->wl, 1);
    orc__iv_serialize(&b, &s->low);
    orc__bv_serialize_into(&b, s->high, s->high_bits);
    orc__select_mcl_serialize_into(&b, s->sel1);
    orc__select_mcl_serialize_into(&b, s->sel0);
    return orc__buf_finish(&b, out, cap);
}
