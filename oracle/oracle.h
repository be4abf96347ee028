/*
 * oracle.h — C API of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * The oracle is a boost-free CPU restatement of the local-score hot path of
 * ninalu/urlearning-cpp.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (liburlgpu.so, the `score` host binary) never links or calls anything here.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/urlearning/).
 *
 * Parity pinning: the reference ships NO tests and NO golden .pss for this
 * path.  The oracle is pinned against (a) oracle/_ref — the
 * reference's own BIC/AD-tree/enumeration/prune sources compiled against shim
 * headers (see oracle/Makefile, oracle/shim/), and (b) the Figure_1/2 golden
 * DAG/MEC matrices (MEC-level).  For cBIC, oracle/_ref/libref_cbic.so compiles the
 * reference's own BIC_OLS.cpp + score_calculator.cpp over a minimal Armadillo / mlpack
 * (oracle/shim_arma): the standardisation, the score formula, the acceptance recursion
 * as written and the store loop are PINNED to that compiled code (tests/test_ref_pin_cbic.py:
 * key sets and values bit-equal).  The floating-point arithmetic INSIDE mlpack/Armadillo
 * (absent, unpinned versions) is restated from their published algorithms on both sides:
 * that part stays "parity unpinned" except through (b).
 */
#ifndef URL_ORACLE_H
#define URL_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---------- input tables (record_file.h:39-54, record.h:35-39, variable.h:43-64) ---------- */
typedef struct orc_table orc_table;
/* returns NULL on error; orc_last_error() has the message */
orc_table *orc_read_csv(const char *path, char delimiter, int has_header);
void orc_table_free(orc_table *t);
int64_t orc_table_n(const orc_table *t);
int orc_table_p(const orc_table *t);
const char *orc_table_name(const orc_table *t, int var);
/* arity = number of distinct value STRINGS in first-appearance order (variable.h:43-48) */
void orc_table_card(const orc_table *t, int32_t *card_out);
/* column-major n*p value indices; returns -1 if some arity > 255 */
int orc_table_codes(const orc_table *t, uint8_t *codes_colmajor);
/* column-major n*p doubles parsed with strtod (mlpack::data::Load, BIC_OLS.cpp:48) */
void orc_table_values(const orc_table *t, double *x_colmajor);
const char *orc_last_error(void);

/* ---------- skeleton (skeleton.cpp:19-105, skeleton.hpp:57-60) ---------- */
/* edges_out: p words (p<=64). returns 1 if initialised from file, 0 if path is NULL/empty (all-ones), <0 on error */
int orc_read_skeleton(const char *path, int p, uint64_t *edges_out);
/* score_main.cpp:145-155: N(v) | U_{j in N(v), j!=v} N(j); initialised==0 -> all_bit_set */
uint64_t orc_two_hop(const uint64_t *edges, int p, int initialised, int v);

/* ---------- enumeration (score_calculator.cpp:54-135, typedefs.h:692-697) ---------- */
/* score_main.cpp:296-304 */
int orc_effective_max_parents(int max_parents_flag, int p, int64_t n_records, int is_bic);
/* masks in the reference's call order: empty set first, then layers 1..max_parents, Gosper order,
 * masks containing v skipped. Returns count (also when masks_out==NULL). */
int64_t orc_enumerate(int v, uint64_t neighbors, int p, int max_parents, uint64_t *masks_out, int64_t cap);

/* ---------- discrete BIC (bic_scoring_function.cpp, log_likelihood_calculator.{h,cpp}) ---------- */
/* cells of the dense contingency table of (v, parents): r_v * prod r_pa */
int64_t orc_bic_cells(const int32_t *card, int p, int v, uint64_t parents);
/* counts[idx], idx = x_v + r_v * paIdx, paIdx mixed radix with the lowest-index parent least
 * significant (log_likelihood_calculator.cpp:61-73) */
int orc_bic_counts(const uint8_t *codes, int64_t n, int p, const int32_t *card, int v, uint64_t parents,
                   int32_t *counts_out);
/* mode 0: Q4 rule: float table ilogi, exact FP64 sum, one rounding to float32, then
 *         score -= tVal*base in float32 (bic_scoring_function.cpp:73).
 * mode 1: literal float32 running sum in contingency-tree DFS order (log_likelihood_calculator.cpp:41-77).
 * ll_out (optional): the pre-rounding FP64 log-likelihood of mode 0. */
int orc_bic_score(const uint8_t *codes, int64_t n, int p, const int32_t *card, int v, uint64_t parents,
                  int mode, float *score_out, double *ll_out);
/* many sets, `threads` host threads (sets striped) */
int orc_bic_score_many(const uint8_t *codes, int64_t n, int p, const int32_t *card, int v,
                       const uint64_t *parents, int64_t n_sets, int mode, int threads, float *scores_out);

/* ---------- discrete fNML (fnml_scoring_function.{h,cpp}) ---------- */
/* one row of getRegretCache: out[N] = (float)log(reg(N, r)), N = 0..n_max */
void orc_log_regret(int64_t n_max, int r, float *out);
/* mode 0: exact-integer contract (float tables on the 2^-23 grid, one final rounding); mode 1: literal float32 sums */
int orc_fnml_score_many(const uint8_t *codes, int64_t n, int p, const int32_t *card, int v,
                        const uint64_t *parents, int64_t n_sets, int mode, int threads, float *scores_out);

/* ---------- discrete BDeu (bdeu_scoring_function.cpp:25-168) ---------- */
/* mode 0: FP64 brackets on the 2^-30 grid, exact sum, one final rounding; mode 1: literal float32 running sum */
int orc_bdeu_score_many(const uint8_t *codes, int64_t n, int p, const int32_t *card, int v, float ess,
                        const uint64_t *parents, int64_t n_sets, int mode, int threads, float *scores_out);

/* ---------- continuous cBIC (BIC_OLS.cpp) ---------- */
/* BIC_OLS.cpp:66-80: centre, divide by sample std (N-1). z_out column-major n*p */
void orc_standardise(const double *x, int64_t n, int p, double *z_out);
/* G = Z^T Z, FP64 (p*p, row-major == symmetric) */
void orc_gram(const double *z, int64_t n, int p, double *g_out);
/* BIC_OLS.cpp:277-389 + mlpack 3.x LinearRegression Train/ComputeError: the_score (double, before negation) */
double orc_cbic_the_score_residual(const double *z, int64_t n, int p, int v, uint64_t parents, double lambda);
/* same quantity from the Gram: RSS = G_vv - g^T G_SS^-1 g via Cholesky */
double orc_cbic_the_score_gram(const double *g, int64_t n, int p, int v, uint64_t parents, double lambda);
/* Store rule (BIC_OLS.cpp:174-276 + score_calculator.cpp:57-61,111-113) applied to float the_scores given in
 * orc_enumerate order. accept_mode 0 = clean (max over cached proper subsets via recursion without the
 * uvec bug), 1 = literal-zero (as written, Armadillo>=10.5 zero fill).
 * stored_out[i] in {0,1}; value_out[i] = stored score (-the_score). */
int orc_cbic_accept(int v, int p, const uint64_t *masks, const float *the_scores, int64_t n_sets,
                    int accept_mode, uint8_t *stored_out, float *value_out);

/* ---------- prune (score_calculator.cpp:137-197) ---------- */
/* literal restatement: sort (score desc, |d|<=2eps -> mask asc), O(m^2) subset scan. keep_out[i] in {0,1} */
int orc_prune(const uint64_t *masks, const float *scores, int64_t m, int highest_completed_layer,
              uint8_t *keep_out);

/* ---------- .pss (score_main.cpp:173-203,383-402; score_cache.cpp:55-162) ---------- */
typedef struct orc_options {
    const char *input;      /* positional 1 */
    const char *output;     /* positional 2 */
    const char *skeleton;   /* -k, may be NULL */
    const char *function;   /* -f: "BIC" | "fNML" | "BDeu" | "cBIC" */
    char delimiter;         /* -d */
    int has_header;         /* -s */
    int max_parents;        /* -p (0 = no limit) */
    double lambda;          /* -l */
    int threads;            /* -t */
    int prune;              /* opt-in restatement of the commented-out call score_main.cpp:166-171 */
    int accept_mode;        /* cBIC: 0 clean, 1 literal-zero */
    int bic_mode;           /* 0 Q4 rule, 1 literal float32 */
    int cbic_from_gram;     /* 0 residual form (reference), 1 Gram/Cholesky form */
    float ess;              /* -e: BDeu equivalent sample size (0 -> the default 1) */
} orc_options;
/* whole `score` run, canonical line order (|S|, mask). returns number of scores written, <0 on error */
int64_t orc_score_file(const orc_options *opt);

#ifdef __cplusplus
}
#endif
#endif
