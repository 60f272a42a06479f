/* oracle/oracle_priv.h — helpers shared between the oracle translation units (test infrastructure). */
#ifndef SDSL_B200_ORACLE_PRIV_H
#define SDSL_B200_ORACLE_PRIV_H
#include "oracle.h"

uint64_t orc__lo_set(uint32_t k);
void orc__write_int(uint64_t *d, uint64_t bitpos, uint64_t x, uint8_t len);
void orc__buf_put(orc_buf *b, const void *src, uint64_t n);
void orc__buf_u64(orc_buf *b, uint64_t x);
uint64_t orc__buf_finish(orc_buf *b, uint8_t *out, uint64_t cap);
void orc__iv_init(orc_iv *v, uint64_t size, uint8_t width);
void orc__iv_set(orc_iv *v, uint64_t i, uint64_t x);
void orc__iv_serialize(orc_buf *b, const orc_iv *v);
/* append the serialised forms used inside composite structures */
void orc__bv_serialize_into(orc_buf *b, const uint64_t *w, uint64_t nbits);
void orc__rank_v_serialize_into(orc_buf *b, const uint64_t *table, uint64_t nbits);
void orc__select_mcl_serialize_into(orc_buf *b, const orc_selmcl *s);
void orc__wt_huff_serialize_into(orc_buf *b, const orc_wt_huff *w);
#endif
