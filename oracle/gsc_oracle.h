/*
 * gsc_oracle.h -- CPU restatement of the SoundChunks encoder hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and there only as the checker.
 *
 * PARITY UNPINNED: the reference (bravesoftdz/soundchunks) ships no golden
 * vectors, known-answer tests or fixtures for this path, and it cannot be
 * built or run here (FreePascal + Windows-only DLLs without source, see
 * DESIGN.md).  This file restates, function by function, the algorithm of
 *   encoder/encoder.lpr   (cited as  enc:LINE)
 *   decoder/decoder.lpr   (cited as  dec:LINE)
 *   encoder/extern.pas    (cited as  ext:LINE)
 * plus the behaviour of the two third-party binaries the encoder binds:
 *   yakmo  (encoder/yakmo_single.dll, no version pin; behaviour recovered by
 *           disassembly of init() RVA 0x16f0 / run() RVA 0x21d0)
 *   ANN 1.1.2 float build (encoder/ANN.dll) -- exact (eps = 0) k-NN, squared
 *           L2 accumulated left-to-right in float, no FMA.
 *
 * Data layout shared with the CUDA library (include/gsc_cuda.h):
 *   pcm        int16 planar  [C][S]            (enc:1143-1145 de-interleave)
 *   chunk n    = i*C + ch  (enc:475-484), cs samples each
 *   attr byte  = bit0 Reversed | bit1 Negative (same code as the KNNFit
 *                variant number neg*2+rev, enc:960-961)
 *   features   float [N][2*cs]                 (enc:802-806)
 */
#ifndef GSC_ORACLE_H
#define GSC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- scalar sample functions (enc:1638-1698) ---- */
double  gsc_ref_log_cr(double x);   /* the shared correctly-rounded ln (soundchunks_b200/csrc/gsc_log.h) */
double  gsc_ref_float_sample(int16_t s);
int16_t gsc_ref_make16(double smp);
int16_t gsc_ref_quant(double smp, int bits, int atten, int neg, double law);
double  gsc_ref_dequant(int16_t q, int bits, int atten, int neg, double law);
int     gsc_ref_attenuation(int cs, const double *x, double law);

/* ---- per chunk (enc:349-397, 1700-1716, 258-322) ---- */
void gsc_ref_chunk_attrs(int cs, const double *x, double law,
                         int *atten, int *neg, int *rev);
void gsc_ref_chunk_features(int cs, const double *x, int neg, int rev,
                            double *f /* 2*cs */);

/* ---- per frame ---- */
/* enc:566-605.  pcm: planar, row stride `stride` samples. v_out[64] optional. */
int gsc_ref_find_attenuation_divider(const int16_t *pcm, int64_t stride,
                                     int C, int S, int cs, int bits,
                                     double *v_out);

/* enc:467-485 + 802-806: all chunks of a frame.
 * raw  double[N][cs]; attr u8[N]; atten u8[N]; feat float[N][2cs];
 * dst int16[N][cs] (may be NULL). Returns N. */
int gsc_ref_make_chunks(const int16_t *pcm, int64_t stride, int C, int S,
                        int cs, int bits, int divider,
                        double *raw, uint8_t *attr, uint8_t *atten,
                        float *feat, int16_t *dst);

/* yakmo as called at enc:824-828 (k-means++ / random seeding, `max_iter`
 * Lloyd rounds in yakmo's run() sense: 0 = one mean update + one
 * reassignment).  seeds_out[K] (optional) receives the chosen point ids. */
void gsc_ref_yakmo(const float *X, int N, int D, int K, int init_type,
                   int max_iter, float *centroids, int32_t *labels,
                   int32_t *seeds_out);

/* enc:699-765 with brute-force exact NN on the live centroids (lowest index
 * on exact ties).  Returns the number of passes; *err_out = last err. */
int gsc_ref_knn_scan_reduce(const float *X, int N, int D, float *centroids,
                            int K, int precision, int max_passes,
                            int32_t *labels, double *err_out);

/* Mini-batched variant of the same rule: assign `batch`
 * points against frozen centroids, then apply the updates in point order. */
int gsc_ref_knn_scan_reduce_batched(const float *X, int N, int D,
                                    float *centroids, int K, int precision,
                                    int max_passes, int batch,
                                    int32_t *labels, double *err_out);

/* enc:699-765 searched the way the binary searches it: ANN 1.1.2 kd-tree (bucket 1, ANN_KD_STD) over the live
 * centroid rows, rebuilt every pass.  stats (optional): [0] points visited, [1] queries whose answer is not the
 * exact nearest centroid (lowest index). */
int gsc_ref_knn_scan_reduce_kdtree(const float *X, int N, int D, float *centroids, int K, int precision,
                                   int max_passes, int32_t *labels, double *err_out, long *stats);

/* enc:915-965 through the same kd-tree (64-NN, epsilon band); best == gsc_ref_knnfit's wherever overfull == 0. */
void gsc_ref_knnfit_kdtree(const int16_t *dict, const uint8_t *datten, int R, int cs, int bits, int divider,
                           const double *raw, int N, int32_t *best, int32_t *use);

/* Plain batch Lloyd: `iters` x (assign to nearest, lowest index on ties;
 * mean in float, point order).  Empty clusters keep their centroid. */
void gsc_ref_lloyd(const float *X, int N, int D, float *centroids, int K,
                   int iters, int32_t *labels);

/* Exact NN labels only (ANN-style distance). dist_out optional. */
void gsc_ref_assign(const float *X, int N, int D, const float *centroids,
                    int K, int32_t *labels, float *dist_out);

/* enc:843-889: class means in the sample domain, population sort (FPC
 * quicksort), dictionary quantisation.
 *   means  float[K][cs]   in sorted (dictionary) order
 *   order  int32[K]       order[i] = cluster id at dictionary position i
 *   counts int32[K]       population at dictionary position i
 *   dict   int16[K][cs], datten u8[K], dattr u8[K] (bit1 neg, bit0 rev)
 *   entry  int32[N]       dictionary position of each chunk (CIInv[label]) */
void gsc_ref_build_dictionary(const int32_t *labels, const double *raw,
                              const uint8_t *attr, int N, int cs, int K,
                              int bits, int divider, float *means,
                              int32_t *order, int32_t *counts, int16_t *dict,
                              uint8_t *datten, uint8_t *dattr,
                              int32_t *entry);

/* enc:891-912 passthrough (N <= K or Precision = 0). */
void gsc_ref_passthrough_dictionary(const double *raw, int N, int cs, int bits,
                                    int divider, int16_t *dict,
                                    uint8_t *datten, uint8_t *dattr);

/* enc:915-965.  Returns the Single-typed epsilon.
 *   best     int32[N]  variant row = entry*4 + neg*2 + rev
 *   use      int32[R]
 *   band     int32[N]  (optional) rows inside the epsilon band (all 4R rows)
 *   best_all int32[N]  (optional) lowest in-band index over ALL rows, i.e.
 *                      without ANN's 64-candidate truncation
 *   dbl_diff int32[N]  (optional) 1 where a Double-typed band test would
 *                      have picked another row */
float gsc_ref_knnfit(const int16_t *dict, const uint8_t *datten, int R, int cs,
                     int bits, int divider, const double *raw, int N,
                     int32_t *best, int32_t *use, int32_t *band,
                     int32_t *best_all, int32_t *dbl_diff);

/* enc:930-938: the 4R x cs Single search set. */
void gsc_ref_knnfit_variants(const int16_t *dict, const uint8_t *datten, int R,
                             int cs, int bits, int divider, float *V);

/* enc:970-977: prune use==0, sort by use desc (FPC quicksort), renumber.
 * remap int32[R]: old position -> new index or -1.  Returns new R. */
int gsc_ref_finalize_dictionary(const int32_t *use, int R, int32_t *remap,
                                int32_t *new_order);

/* FPC fgl TFPSList.QuickSort on an index permutation, comparing keys
 * descending (CompareValue(Item2.key, Item1.key)). perm in/out. */
void gsc_ref_fpc_sort_desc(const int32_t *keys, int32_t *perm, int n);

/* ---- whole frame / whole file ---- */
typedef struct gsc_ref_params {
    int chunk_size;        /* -cs  default 4     enc:1496 */
    int chunk_bit_depth;   /* -cbd default 8     enc:1495 */
    int chunks_per_frame;  /* -cpf default 4096  enc:1505 */
    int precision;         /* -pr  default 3     enc:1503 */
    int max_passes;        /* CMaxIterations = 100, enc:703 */
    int kmeans_mode;       /* 0 online, exact search on the live centroids (the library's contract); 1 lloyd;
                              2 online batched; 3 online through an ANN-1.1.2-style kd-tree rebuilt per pass
                              (planes go stale while the pass moves the rows: what the shipped binary does) */
    int lloyd_iters;       /* mode 1 */
    int batch;             /* mode 2 */
    double frame_length_ms;/* -fl  default 4000  enc:1501 */
    double vfr;            /* -vfr default 1.0   enc:1499 */
    int band_all;          /* KNNFit band rule: 0 = among the 64 nearest rows (ANN's bucket, enc:917,952;
                              ties in index order where ANN's visit order is not reproducible),
                              1 = among ALL rows inside the epsilon band (what libgsc_cuda computes; the
                              two agree whenever overfull == 0) */
    int reserved;
} gsc_ref_params;

void gsc_ref_default_params(gsc_ref_params *p);

typedef struct gsc_ref_frame_out {
    int N, R, divider, passes;
    double err;
    int16_t *dict;       /* [R][cs]  final order */
    uint8_t *datten;     /* [R] */
    int32_t *index;      /* [N] final dictionary index per chunk */
    uint8_t *attr;       /* [N] bit1 neg bit0 rev (after KNNFit) */
    int overfull;        /* queries with > 64 rows in the epsilon band */
} gsc_ref_frame_out;

/* enc:1433-1447 DoFrame for one frame. Caller frees with gsc_ref_free_frame. */
int  gsc_ref_encode_frame(const int16_t *pcm, int64_t stride, int C, int S,
                          const gsc_ref_params *p, gsc_ref_frame_out *out);
void gsc_ref_free_frame(gsc_ref_frame_out *out);

/* enc:1294-1429 frame cut (pass 2). pcm planar [C][S_padded]. starts[] gets
 * the first sample of each frame; returns the frame count (<= max_frames). */
int gsc_ref_plan_frames(const int16_t *pcm, int64_t stride, int C,
                        int64_t S_padded, int sample_rate,
                        const gsc_ref_params *p, int64_t *starts,
                        int max_frames);

/* enc:980-1107 one frame of .gsc.  Returns bytes written (buf may be NULL to
 * size). */
int64_t gsc_ref_write_frame(const gsc_ref_frame_out *f, int C, int cs, int bits,
                            int sample_rate, uint8_t *buf, int64_t cap);

/* dec:37-220.  Decodes a whole .gsc stream to interleaved int16.
 * Returns samples per channel written, or -1. */
int64_t gsc_ref_decode(const uint8_t *gsc, int64_t len, int16_t *out,
                       int64_t cap_samples, int *channels, int *sample_rate);

/* enc:487-522 + 1518-1582: encoder-side reconstruction of one frame (planar
 * int16 out[C][S]). */
void gsc_ref_reconstruct_frame(const gsc_ref_frame_out *f, int C, int S, int cs,
                               int bits, int16_t *out, int64_t stride);

/* enc:1862-1880 / 1816-1827. */
double gsc_ref_psy_a_delta(const int16_t *a, const int16_t *b, int64_t n);
double gsc_ref_snr_db(const int16_t *ref, const int16_t *tst, int64_t n);

#ifdef __cplusplus
}
#endif
#endif
