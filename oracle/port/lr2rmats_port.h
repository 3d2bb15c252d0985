/* oracle/port/lr2rmats_port.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the lr2rmats hot path (see lr2rmats_port.c).  It produces
 * results in the same structure-of-arrays layout as the product C ABI
 * (include/lr2rmats_b200.h) so tests can compare array by array.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this; the product never does.
 *
 * Parity pin: validated byte-for-byte against the compiled reference binary
 * (oracle/_ref/lr2rmats) through the text emitters on the fixtures of SURVEY
 * App. C and on seeded synthetic sets -- see tests/test_oracle_pin.py and
 * tests/golden/.
 */
#ifndef LR2RMATS_PORT_H
#define LR2RMATS_PORT_H
#include "../../include/lr2rmats_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* All orc_* calls malloc the arrays they return inside the result structs;
 * release with the matching orc_free_*. */
int  orc_filter(const lrb_batch *b, const lrb_anno *rm, const lrb_filter_params *p, lrb_filter_result *out);
void orc_free_filter(lrb_filter_result *r);

/* sel==NULL: every record */
int  orc_bam2gtf(const lrb_batch *b, const uint32_t *sel, int64_t n_sel, const lrb_exon_params *p, lrb_exon_result *out);
void orc_free_exon(lrb_exon_result *r);

/* chains: read-derived transcripts (from orc_bam2gtf or -m g input) */
int  orc_update(const lrb_exon_result *chains, const lrb_anno *anno, const lrb_sj *sj,
                const lrb_update_params *p, lrb_update_result *out);
void orc_free_update(lrb_update_result *r);

int  orc_unique(const lrb_exon_result *chains, const lrb_update_params *p, lrb_unique_result *out);
void orc_free_unique(lrb_unique_result *r);

/* bam2sj_core (parse_bam.c:896-924): distinct junctions with uniq / multi counts, in the order the reference's insertion leaves them */
int  orc_bam2sj(const lrb_batch *b, const uint8_t *is_uniq, const lrb_sj_params *p, lrb_sj *out);
void orc_free_sj(lrb_sj *r);

#ifdef __cplusplus
}
#endif
#endif
