// gsc_device.cuh -- device-side arithmetic shared by every kernel.
//
// Bit-exactness rules (the whole library is compiled with -fmad=false):
//   * `a*b+c` is never contracted; where an FFMA is wanted it is written fmaf().
//   * Double paths follow enc:1638-1698 operation by operation (IEEE div/sqrt).
//   * round() of FreePascal = half-to-even = __double2ll_rn.
// enc:L = /root/reference/encoder/encoder.lpr line L.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define GSC_MAX_ATT 15        // enc:14 CMaxAttenuation
#define GSC_MAX_K 4096        // enc:15 CMaxChunksPerFrame
#define GSC_BUCKET 64         // enc:917 CBucketSize
#define GSC_MAX_CS 8          // largest ChunkSize the kernels are instantiated for

// Per-frame descriptor, one per frame of a batch, in device memory.
struct GscFrame {
    long long pcm_off;    // offset (samples) of channel 0, sample 0 in the batch PCM buffer
    long long stride;     // samples between channel rows
    long long chunk_off;  // prefix sum of N over the batch
    int C, S, cc, N;      // channels, samples per channel, chunks per channel, N = cc*C
    int K;                // clusters (0 => passthrough frame, enc:891-912)
    int R;                // dictionary entries before pruning (K, or N when passthrough)
    int slot;             // index of the frame in per-frame arrays
    int pad;
};

// ---- sample conversion (enc:1643-1646) --------------------------------------
__device__ __forceinline__ double gsc_sample(short s) { return (double)s / 32767.0; }

// Attenuation coefficient table for one Law (enc:1654-1656, 1692-1696):
// T[0] = 1.0 (+ 0*Law), T[r] = T[r-1] + r*Law, r = 1..16.
struct GscLaw {
    double T[GSC_MAX_ATT + 2];
    __device__ __forceinline__ void init(double law) {
        T[0] = 1.0 + 0.0 * law;
#pragma unroll
        for (int r = 1; r <= GSC_MAX_ATT + 1; ++r) T[r] = T[r - 1] + (double)r * law;
    }
};

// enc:1682-1698 ComputeAttenuation given hi = max ceil(|x*32767|)
__device__ __forceinline__ int gsc_attenuation(int hi, const GscLaw &L) {
    int r = 0;
    do {
        ++r;
    } while (!(((double)hi * L.T[r] > 32767.0) || (r > GSC_MAX_ATT)));
    return r - 1;
}
__device__ __forceinline__ int gsc_hi(double x) { return (int)ceil(fabs(x * 32767.0)); }

// enc:1648-1663 makeOutputSample
__device__ __forceinline__ short gsc_quant(double smp, int obd, double coeff, bool neg) {
    long long r = __double2ll_rn(smp * (double)obd * coeff);
    short s = (short)r;
    if (neg) s = (short)(-s);
    if (s < -obd + 1) s = (short)(-obd + 1);
    if (s > obd - 1) s = (short)(obd - 1);
    return s;
}
// enc:1665-1680 makeFloatSample(5 args)
__device__ __forceinline__ double gsc_dequant(short q, int obd, double coeff, bool neg) {
    short s = q;
    if (neg) s = (short)(-s);
    double r = (double)s / ((double)obd * coeff);
    if (r < -1.0) r = -1.0;
    if (r > 1.0) r = 1.0;
    return r;
}

// enc:365-397 ComputeDstAttributes on cs doubles -> (atten, neg, rev)
template <int CS>
__device__ __forceinline__ void gsc_chunk_attrs(const double (&x)[CS], const GscLaw &L,
                                                int &atten, bool &neg, bool &rev) {
    int hi = 0;
#pragma unroll
    for (int i = 0; i < CS; ++i) { int v = gsc_hi(x[i]); hi = v > hi ? v : hi; }
    atten = gsc_attenuation(hi, L);
    double p1 = 0.0, p2 = 0.0;
#pragma unroll
    for (int i = 0; i < CS; ++i) if (x[i] < 0) p1 -= x[i];
#pragma unroll
    for (int i = 0; i < CS; ++i) if (x[i] > 0) p2 += x[i];
    neg = p1 > p2;
    p1 = 0.0; p2 = 0.0;
#pragma unroll
    for (int i = 0; i < CS / 2; ++i) p1 += fabs(x[i]);
#pragma unroll
    for (int i = CS / 2; i < CS; ++i) p2 += fabs(x[i]);
    rev = p1 > p2;
}

// ANN distance (float, left to right, separate multiply and add): d = 0; for k: t = q_k - p_k; d = d + t*t.
// The subtractions and the squarings of two neighbouring dimensions are independent IEEE operations, so they go through
// the packed FP32 pipe (sub.rn.f32x2 / mul.rn.f32x2 -> FADD2 / FMUL2: two results per issue slot, bit for bit the scalar
// results); the accumulation stays a chain of scalar adds in the reference's order.  (Not add.rn.f32x2: ptxas 12.9
// contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with explicit .rn and -fmad=false, which would round once
// instead of twice; a packed multiply followed by scalar adds is left alone -- checked in the SASS.)
__device__ __forceinline__ unsigned long long gsc_pk2f(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void gsc_upk2f(unsigned long long v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
template <int D>
__device__ __forceinline__ float gsc_ann_dist(const float (&q)[D], const float (&p)[D]) {
    float d = 0.0f;
    if (D % 2 == 0) {
#pragma unroll
        for (int k = 0; k < D / 2; ++k) {
            unsigned long long t, m;
            asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(gsc_pk2f(q[2 * k], q[2 * k + 1])), "l"(gsc_pk2f(p[2 * k], p[2 * k + 1])));
            asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(m) : "l"(t), "l"(t));
            float m0, m1;
            gsc_upk2f(m, m0, m1);
            d = (k == 0) ? m0 : d + m0;      // 0 + m0 == m0 bit for bit: a square is +0, positive or NaN, never -0
            d = d + m1;
        }
    } else {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            float t = q[k] - p[k];
            float m = t * t;
            d = d + m;
        }
    }
    return d;
}

// The same distance for TWO codebook rows at once: p2[k] holds (row A's k-th coordinate, row B's) as a packed pair, so the
// subtraction, the squaring AND the ordered accumulation of both rows go through the packed pipe: 3 instructions per
// dimension for two rows.  The accumulation d = d + m is issued as fma(m, 1, d): the product m * 1 is exact, so the one
// rounding of the fma is the rounding of the addition, bit for bit.  `ones` must be (1.0f, 1.0f) read at RUN time (a
// kernel parameter): with a literal 1 ptxas rewrites fma(m,1,d) to an add and then contracts the preceding multiply
// into it (FFMA2 of t*t+d: one rounding where the reference has two) -- seen in the SASS, hence the detour.
template <int D>
__device__ __forceinline__ void gsc_ann_dist2(const float (&q)[D], const unsigned long long (&p2)[D], unsigned long long ones,
                                              float &dA, float &dB) {
    unsigned long long d2 = 0ull;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        unsigned long long t, m;
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(gsc_pk2f(q[k], q[k])), "l"(p2[k]));
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(m) : "l"(t), "l"(t));
        if (k == 0) d2 = m;                   // 0 + m == m: a square is +0, positive or NaN
        else asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d2) : "l"(m), "l"(ones), "l"(d2));
    }
    gsc_upk2f(d2, dA, dB);
}

// yakmo distance (init() RVA 0x1dca): d = (cn + pn) + 0; d -= (p_k + p_k) * c_k
template <int D>
__device__ __forceinline__ float gsc_yakmo_dist(const float (&p)[D], float pn,
                                                const float (&c)[D], float cn) {
    float d = cn + pn;
    d = d + 0.0f;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        float t = p[k] + p[k];
        t = t * c[k];
        d = d - t;
    }
    return d;
}

// FreePascal fgl TFPSList.QuickSort on an index permutation, keys descending
// (Compare(Item1, Item2) = CompareValue(Item2.key, Item1.key), enc:775-783).
// Iterative form of the recursion `if L < J then QuickSort(L, J); L := I`.
// One thread; `stack` needs 2*64 ints.
__device__ __forceinline__ int gsc_cmp_inv(const int *keys, int a, int b) {
    int ka = keys[a], kb = keys[b];
    return (kb < ka) ? -1 : (kb > ka) ? 1 : 0;
}
__device__ inline void gsc_fpc_sort_desc(const int *keys, int *items, int n, int *stack) {
    if (n < 2) return;
    // Explicit stack of pending (L, R) outer loops.  The recursion processes
    // (L, J) completely before continuing with (I, R): depth-first, left first.
    int sp = 0;
    int L = 0, R = n - 1;
    for (;;) {
        // body of `repeat ... until I >= R` for the current (L, R)
        int I = L, J = R;
        int P = (int)(((unsigned)L + (unsigned)R) >> 1);
        do {
            int pivot = items[P];
            while (gsc_cmp_inv(keys, pivot, items[I]) > 0) ++I;
            while (gsc_cmp_inv(keys, pivot, items[J]) < 0) --J;
            if (I <= J) {
                int t = items[I]; items[I] = items[J]; items[J] = t;
                if (P == I) P = J; else if (P == J) P = I;
                ++I; --J;
            }
        } while (!(I > J));
        // continuation of this frame is (I, R) if I < R; the recursive call is (L, J) if L < J
        bool rec = L < J;
        bool cont = I < R;
        if (rec) {
            if (cont) { stack[sp++] = I; stack[sp++] = R; }
            R = J;           // L stays
        } else if (cont) {
            L = I;           // R stays
        } else {
            if (sp == 0) break;
            R = stack[--sp]; L = stack[--sp];
        }
    }
}

// Monotone map float -> uint32 (total order, NaN last for positive NaN).
__device__ __forceinline__ unsigned gsc_fkey(float f) {
    unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
