/* oracle/oracle_csa.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See oracle.h.
 *
 * Part 3: csa_wt<wt_huff<>, 32, 64, sa_order_sa_sampling<>, isa_sampling<>, byte_alphabet> (rows a9, a10):
 * construction from a zero-free byte text, backward_search / count, SA access, locate, serialisation.
 * The suffix array itself is built by plain prefix doubling (the reference uses divsufsort; any correct
 * suffix sorter yields the same array, and the serialised index is compared byte for byte with the
 * reference's in tests/test_oracle_csa.py).
 * Citations are relative to /root/reference/include/sdsl/.
 */
#include "oracle_priv.h"

#include <stdlib.h>
#include <string.h>

#define SA_DENS 32u  /* csa_wt.hpp:50 t_dens */
#define ISA_DENS 64u /* csa_wt.hpp:51 t_inv_dens */

/* ---- suffix array by prefix doubling with LSD radix passes on (rank[i], rank[i+k]) ------------ */
static void radix_pass(const uint32_t *in, uint32_t *out, const uint32_t *key, uint64_t n, int shift, uint64_t *cnt)
{
    uint64_t i, s = 0;
    memset(cnt, 0, 65537 * sizeof(uint64_t));
    for (i = 0; i < n; ++i)
        ++cnt[(key[in[i]] >> shift) & 0xFFFF];
    for (i = 0; i < 65536; ++i) {
        uint64_t c = cnt[i];
        cnt[i] = s;
        s += c;
    }
    for (i = 0; i < n; ++i)
        out[cnt[(key[in[i]] >> shift) & 0xFFFF]++] = in[i];
}

/* text t[0..n) where t[n-1] is the unique smallest symbol; returns malloc'ed SA (uint32) */
static uint32_t *suffix_array(const uint8_t *t, uint64_t n)
{
    uint32_t *sa = (uint32_t *)malloc(4 * (n + 1)), *tmp = (uint32_t *)malloc(4 * (n + 1));
    uint32_t *rk = (uint32_t *)malloc(4 * (n + 1)), *rk2 = (uint32_t *)malloc(4 * (n + 1)), *nrk = (uint32_t *)malloc(4 * (n + 1));
    uint64_t *cnt = (uint64_t *)malloc(65537 * sizeof(uint64_t));
    uint64_t i, k;
    for (i = 0; i < n; ++i) {
        sa[i] = (uint32_t)i;
        rk[i] = t[i];
    }
    /* initial order by first symbol */
    radix_pass(sa, tmp, rk, n, 0, cnt);
    memcpy(sa, tmp, 4 * n);
    for (k = 1;; k <<= 1) {
        uint32_t r = 0;
        /* rank of the suffix k positions later; 0 = past the end (ranks are stored +1) */
        if (k == 1)
            for (i = 0; i < n; ++i)
                rk[i] += 1;
        for (i = 0; i < n; ++i)
            rk2[i] = (i + k < n) ? rk[i + k] : 0;
        radix_pass(sa, tmp, rk2, n, 0, cnt);
        radix_pass(tmp, sa, rk2, n, 16, cnt);
        radix_pass(sa, tmp, rk, n, 0, cnt);
        radix_pass(tmp, sa, rk, n, 16, cnt);
        nrk[sa[0]] = r = 1;
        for (i = 1; i < n; ++i) {
            if (rk[sa[i]] != rk[sa[i - 1]] || rk2[sa[i]] != rk2[sa[i - 1]])
                ++r;
            nrk[sa[i]] = r;
        }
        memcpy(rk, nrk, 4 * n);
        if (r == n)
            break;
    }
    free(tmp);
    free(rk);
    free(rk2);
    free(nrk);
    free(cnt);
    return sa;
}

/* csa_wt.hpp:323-355 via construct.hpp:127-193: text + 0 sentinel -> SA -> BWT -> alphabet, samples, WT */
orc_csa *orc_csa_build(const uint8_t *text, uint64_t len)
{
    orc_csa *c = (orc_csa *)calloc(1, sizeof(*c));
    uint64_t n = len + 1, i;
    uint8_t *t = (uint8_t *)malloc(n), *bwt = (uint8_t *)malloc(n);
    uint32_t *sa;
    uint64_t cnt[256];
    memcpy(t, text, len);
    t[len] = 0; /* construct.hpp:47-52 append_zero_symbol */
    sa = suffix_array(t, n);
    for (i = 0; i < n; ++i) /* construct_bwt.hpp:53-60 */
        bwt[i] = sa[i] ? t[sa[i] - 1] : t[n - 1];
    c->n = n;
    /* byte_alphabet, csa_alphabet_strategy.hpp:175-212 */
    memset(cnt, 0, sizeof(cnt));
    for (i = 0; i < n; ++i)
        ++cnt[bwt[i]];
    c->sigma = 0;
    for (i = 0; i < 256; ++i)
        if (cnt[i]) {
            c->char2comp[i] = (uint8_t)c->sigma;
            c->comp2char[c->sigma] = (uint8_t)i;
            c->C[c->sigma + 1] = cnt[i];
            ++c->sigma;
        }
    c->C[0] = 0;
    for (i = 1; i <= c->sigma; ++i)
        c->C[i] += c->C[i - 1];
    /* _sa_order_sampling, csa_sampling_strategy.hpp:98-115: SA[0], SA[32], ... ; width hi(n)+1 */
    orc__iv_init(&c->sa_sample, (n + SA_DENS - 1) / SA_DENS, (uint8_t)(orc_hi(n) + 1));
    for (i = 0; i < n; i += SA_DENS)
        orc__iv_set(&c->sa_sample, i / SA_DENS, sa[i]);
    /* _isa_sampling, csa_sampling_strategy.hpp:758-779: isa_sample[SA[i]/64] = i when SA[i] % 64 == 0 */
    orc__iv_init(&c->isa_sample, (n - 1) / ISA_DENS + 1, (uint8_t)(orc_hi(n) + 1));
    for (i = 0; i < n; ++i)
        if (sa[i] % ISA_DENS == 0)
            orc__iv_set(&c->isa_sample, sa[i] / ISA_DENS, i);
    c->wt = orc_wt_huff_build(bwt, n);
    free(sa);
    free(t);
    free(bwt);
    return c;
}

void orc_csa_free(orc_csa *c)
{
    if (!c)
        return;
    orc_wt_huff_free(c->wt);
    orc_iv_free(&c->sa_sample);
    orc_iv_free(&c->isa_sample);
    free(c);
}

/* suffix_array_algorithm.hpp:166-201 (one character) */
static uint64_t backward_step(const orc_csa *c, uint64_t l, uint64_t r, uint8_t ch, uint64_t *lr, uint64_t *rr)
{
    uint64_t cc = c->char2comp[ch];
    if (cc == 0 && ch > 0) {
        *lr = 1;
        *rr = 0;
    } else {
        uint64_t cb = c->C[cc];
        if (l == 0 && r + 1 == c->n) {
            *lr = cb;
            *rr = c->C[cc + 1] - 1;
        } else {
            *lr = cb + orc_wt_huff_rank(c->wt, l, ch);
            *rr = cb + orc_wt_huff_rank(c->wt, r + 1, ch) - 1;
        }
    }
    return *rr + 1 - *lr;
}

/* suffix_array_algorithm.hpp:227-248 (pattern) and :463-471 (count) */
uint64_t orc_csa_count(const orc_csa *c, const uint8_t *pat, uint64_t m, uint64_t *l_out, uint64_t *r_out)
{
    uint64_t l = 0, r = c->n - 1, it = m;
    if (m > c->n) {
        if (l_out)
            *l_out = 0;
        if (r_out)
            *r_out = 0;
        return 0;
    }
    while (it > 0 && r + 1 - l > 0) {
        --it;
        backward_step(c, l, r, pat[it], &l, &r);
    }
    if (l_out)
        *l_out = l;
    if (r_out)
        *r_out = r;
    return r + 1 - l;
}

/* csa_wt.hpp:363-381 with the LF step of suffix_array_helper.hpp:346-360 */
uint64_t orc_csa_sa(const orc_csa *c, uint64_t i)
{
    uint64_t off = 0, v;
    while (i % SA_DENS != 0) {
        uint64_t sym, j = orc_wt_huff_inverse_select(c->wt, i, &sym);
        i = c->C[c->char2comp[sym]] + j;
        ++off;
    }
    v = orc_iv_get(&c->sa_sample, i / SA_DENS) + off;
    return v < c->n ? v : v - c->n;
}

void orc_csa_count_batch(const orc_csa *c, const uint8_t *pats, const uint64_t *off, uint64_t n, uint64_t *cnt, uint64_t *l_out)
{
    uint64_t k;
    for (k = 0; k < n; ++k)
        cnt[k] = orc_csa_count(c, pats + off[k], off[k + 1] - off[k], l_out ? &l_out[k] : NULL, NULL);
}

/* suffix_array_algorithm.hpp:534-550: occurrences in suffix-array order.  occ_off must already hold the
 * exclusive prefix sums of the counts. */
void orc_csa_locate_batch(const orc_csa *c, const uint8_t *pats, const uint64_t *off, uint64_t n, const uint64_t *occ_off, uint64_t *occ)
{
    uint64_t k, j;
    for (k = 0; k < n; ++k) {
        uint64_t l, r, cnt = orc_csa_count(c, pats + off[k], off[k + 1] - off[k], &l, &r);
        for (j = 0; j < cnt; ++j)
            occ[occ_off[k] + j] = orc_csa_sa(c, l + j);
    }
}

void orc_csa_sa_batch(const orc_csa *c, const uint64_t *i, uint64_t n, uint64_t *out)
{
    uint64_t k;
    for (k = 0; k < n; ++k)
        out[k] = orc_csa_sa(c, i[k]);
}

/* ISA[i] from the samples (suffix_array_helper.hpp:519-537, csa_sampling_strategy.hpp:795-800 sample_qeq) */
uint64_t orc_csa_isa(const orc_csa *c, uint64_t i)
{
    uint64_t ci = (i / ISA_DENS + 1) % c->isa_sample.size, pos = ci * ISA_DENS, r = orc_iv_get(&c->isa_sample, ci), steps;
    steps = pos < i ? pos + c->n - i : pos - i;
    while (steps--) {
        uint64_t sym, j = orc_wt_huff_inverse_select(c->wt, r, &sym);
        r = c->C[c->char2comp[sym]] + j;
    }
    return r;
}

/* extract(csa, begin, end) for the LF-based CSA (suffix_array_algorithm.hpp:590-610): text[begin..end] inclusive */
void orc_csa_extract(const orc_csa *c, uint64_t begin, uint64_t end, uint8_t *out)
{
    uint64_t steps = end - begin + 1, order = orc_csa_isa(c, end), k;
    /* first_row_symbol(order): the symbol whose C-bucket contains `order` (suffix_array_helper.hpp) */
    for (k = 0; k + 1 < c->sigma && c->C[k + 1] <= order; ++k)
        ;
    out[--steps] = c->comp2char[k];
    while (steps != 0) {
        uint64_t sym, j = orc_wt_huff_inverse_select(c->wt, order, &sym);
        order = c->C[c->char2comp[sym]] + j;
        out[--steps] = (uint8_t)sym;
    }
}

void orc_csa_extract_batch(const orc_csa *c, const uint64_t *begin, const uint64_t *end, uint64_t n, const uint64_t *out_off, uint8_t *out)
{
    uint64_t k;
    for (k = 0; k < n; ++k)
        orc_csa_extract(c, begin[k], end[k], out + out_off[k]);
}

/* csa_wt.hpp:389-402 + csa_alphabet_strategy.hpp:258-268 */
uint64_t orc_csa_serialize(const orc_csa *c, uint8_t *out, uint64_t cap)
{
    orc_buf b = {0, 0, 0};
    orc_iv v;
    uint64_t k;
    orc__wt_huff_serialize_into(&b, c->wt);
    orc__iv_serialize(&b, &c->sa_sample);
    orc__iv_serialize(&b, &c->isa_sample);
    orc__iv_init(&v, 256, 8);
    for (k = 0; k < 256; ++k)
        orc__iv_set(&v, k, c->char2comp[k]);
    orc__iv_serialize(&b, &v);
    orc_iv_free(&v);
    orc__iv_init(&v, c->sigma, 8);
    for (k = 0; k < c->sigma; ++k)
        orc__iv_set(&v, k, c->comp2char[k]);
    orc__iv_serialize(&b, &v);
    orc_iv_free(&v);
    orc__iv_init(&v, (uint64_t)c->sigma + 1, 64);
    for (k = 0; k <= c->sigma; ++k)
        orc__iv_set(&v, k, c->C[k]);
    orc__iv_serialize(&b, &v);
    orc_iv_free(&v);
    orc__buf_put(&b, &c->sigma, 2);
    return orc__buf_finish(&b, out, cap);
}
