// gsc_plan.cuh -- SURVEY.md 8(f1): the frame planner's power scan and boundary selection (TEncoder.PrepareFrames,
// enc:1374-1425, with the int16 -> Double staging of enc:1282-1285) on the device, bit-exact.
//
// What the reference computes, all in Double and all as SEQUENTIAL sums over the whole file:
//   A     = sum_{ch} sum_i x^2,  x = s / 32767.0           avg = sqrt(A / (S * C))                  enc:1376-1386
//   T     = sum_i t_i,  t_i = 1 - lerp(avg, smp_i, vfr),  smp_i = sqrt((sum_ch x^2) / C)            enc:1388-1398
//   per   = T / frameCount;  cur += t_i;  a frame starts at the first i (multiple of the block size) with cur >= per,
//           cur := 0                                                                                 enc:1402-1423
// Double addition is not associative: a tree sum gives other last bits, other `avg`, other boundaries.  The sums are
// evaluated EXACTLY in parallel with the parity-function scheme of the seeding kernel (gsc_seed.cuh) carried over
// to Double: inside one binade the running sum is an integer S (2^52 <= S < 2^53) times the ulp, an element acts
// on S as a function parity -> increment (round-half-even only looks at the parity), functions compose
// associatively.  Per 2048-element window: a tree sum (for an approximate prefix, which predicts the exponent of
// the running sum at the window's start), then a summary (f0, f1, bounds) under that exponent; one warp chains the
// windows, taking a summary when the actual exponent matches and the bounds keep every partial sum inside the
// binade, adding the window up element by element otherwise (binade crossings, ~60 per file).
// Boundaries: candidates from the approximate prefix (binary search per frame), then every frame is VERIFIED by one
// warp running the reference's loop exactly from its start; a frame whose exact boundary differs from the candidate
// corrects it and the frames behind it are redone (the host loop in gsc_api.cu; expected zero iterations).
#pragma once
#include "gsc_device.cuh"

#define GSC_PW 2048           // elements per window
#define GSC_PT 256            // threads per summary CTA (8 elements each)
#define GSC_PF_BAD 1
#define GSC_PF_ZERO 2         // every element of the window is +0: the running sum does not change

struct GscWin { long long f0, f1, neg, pos; int e0, flags; long long pad; };   // 48 bytes

__device__ __forceinline__ void gsc_pfd_compose(long long &f0, long long &f1, long long g0, long long g1) {
    const long long h0 = f0 + ((f0 & 1) ? g1 : g0);
    const long long h1 = f1 + (((1 + f1) & 1) ? g1 : g0);
    f0 = h0; f1 = h1;
}
__device__ __forceinline__ void gsc_classify_d(double v, double sc, double huge, long long &I, int &cls) {
    const double av = fabs(v);
    if (!(av < huge)) { I = 1ll << 58; cls = 4; return; }
    const double aq = av * sc;              // exact (power-of-two scaling), < 2^55
    const double fl = floor(aq);
    const double g = aq - fl;               // exact fraction
    const long long ni = (long long)fl;
    const int tie = (g == 0.5) ? 2 : 0;
    if (v >= 0.0) { I = ni; cls = tie | ((g > 0.5) ? 1 : 0); }
    else if (g == 0.0) { I = -ni; cls = 0; }
    else { I = -(ni + 1); cls = tie | ((g < 0.5) ? 1 : 0); }
}
__device__ __forceinline__ long long gsc_pfd_step(long long S, long long I, int cls) {
    const long long t = S + I;
    return t + (long long)((cls & 1) | ((cls >> 1) & (int)(t & 1)));
}
__device__ __forceinline__ double gsc_mkd(int e, long long S) {
    return __longlong_as_double(((long long)e << 52) | (S & 0xfffffffffffffll));
}

// ---- terms -----------------------------------------------------------------------------------------------
// pass 1: term n = x^2 of sample n in (channel-major) order, n = ch * S + i
__global__ void k_plan_terms1(const short *__restrict__ pcm, long long stride, int C, long long S, double *__restrict__ out) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= S * C) return;
    const long long ch = n / S, i = n - ch * S;
    const double x = gsc_sample(pcm[ch * stride + i]);
    out[n] = x * x;
}
// pass 2: t_i = 1 - (avg + (smp_i - avg) * vfr)
__global__ void k_plan_terms2(const short *__restrict__ pcm, long long stride, int C, long long S, const double *__restrict__ avgp,
                              double vfr, double *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const double avg = *avgp;
    double smp = 0.0;
    for (int j = 0; j < C; ++j) { const double x = gsc_sample(pcm[(long long)j * stride + i]); smp += x * x; }
    smp = sqrt(smp / (double)C);
    out[i] = 1.0 - (avg + (smp - avg) * vfr);
}

// ---- approximate prefix: tree sum per window, then an exclusive scan over the windows (one CTA) ------------------
__global__ void __launch_bounds__(GSC_PT) k_plan_winsum(const double *__restrict__ a, long long n, double *__restrict__ wsum) {
    __shared__ double s_w[GSC_PT / 32];
    const long long base = (long long)blockIdx.x * GSC_PW;
    double s = 0.0;
#pragma unroll
    for (int e = 0; e < GSC_PW / GSC_PT; ++e) {
        const long long j = base + (long long)e * GSC_PT + threadIdx.x;
        if (j < n) s += a[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < GSC_PT / 32; ++w) t += s_w[w];
        wsum[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(1024) k_plan_winscan(const double *__restrict__ wsum, long long nw, double *__restrict__ wpre) {
    __shared__ double s_t[1024];
    const long long per = (nw + 1023) / 1024, b0 = (long long)threadIdx.x * per, b1 = min(nw, b0 + per);
    double s = 0.0;
    for (long long w = b0; w < b1; ++w) s += wsum[w];
    s_t[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) { double r = 0.0; for (int t = 0; t < 1024; ++t) { const double v = s_t[t]; s_t[t] = r; r += v; } }
    __syncthreads();
    double r = s_t[threadIdx.x];
    for (long long w = b0; w < b1; ++w) { wpre[w] = r; r += wsum[w]; }
    if (threadIdx.x == 1023) wpre[nw] = r;       // (b1 == nw for the last non-empty stretch)
}

// ---- window summaries under the exponent the approximate prefix predicts ----------------------------------------
__global__ void __launch_bounds__(GSC_PT) k_plan_summaries(const double *__restrict__ a, long long n, const double *__restrict__ wpre,
                                                          GscWin *__restrict__ win) {
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ long long s_f0[GSC_PT / 32], s_f1[GSC_PT / 32], s_neg[GSC_PT / 32], s_pos[GSC_PT / 32];
    __shared__ int s_fl[GSC_PT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long base = (long long)blockIdx.x * GSC_PW + (long long)tid * 8;
    double v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (base + e < n) ? a[base + e] : 0.0;
    const double p = wpre[blockIdx.x];
    const int e0 = (int)((__double_as_longlong(p) >> 52) & 0x7ff);
    const bool valid = p > 0.0 && e0 >= 64 && e0 <= 2040;
    bool bad = !valid, allzero = true;
    long long g0 = 0, g1 = 1, neg = 0, pos = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) { allzero &= (__double_as_longlong(v[e]) == 0ll); bad |= !(fabs(v[e]) < 1.0e300); }
    if (valid) {
        const double sc = __longlong_as_double((long long)(2098 - e0) << 52);      // 1 / ulp = 2^(1075 - e0)
        const double huge = __longlong_as_double((long long)(e0 + 2) << 52);       // 4 * 2^(e0 - 1023)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            long long I; int cls;
            gsc_classify_d(v[e], sc, huge, I, cls);
            bad |= cls >= 4;
            g0 = gsc_pfd_step(g0, I, cls); g1 = gsc_pfd_step(g1, I, cls);
            if (I < 0) neg += -I; else pos += I + 1;
        }
        g1 -= 1;
        neg = min(neg, 1ll << 56); pos = min(pos, 1ll << 56);
    } else { g0 = 0; g1 = 0; }
    long long s0 = g0, s1 = g1;
#pragma unroll
    for (int of = 1; of < 32; of <<= 1) {
        const long long p0 = __shfl_up_sync(FULL, s0, of), p1 = __shfl_up_sync(FULL, s1, of);
        if (lane >= of) { long long q0 = p0, q1 = p1; gsc_pfd_compose(q0, q1, s0, s1); s0 = q0; s1 = q1; }
    }
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) { neg += __shfl_xor_sync(FULL, neg, of); pos += __shfl_xor_sync(FULL, pos, of); }
    const bool wbad = __any_sync(FULL, bad), wzero = __all_sync(FULL, allzero);
    if (lane == 31) { s_f0[warp] = s0; s_f1[warp] = s1; }
    if (lane == 0) { s_neg[warp] = min(neg, 1ll << 56); s_pos[warp] = min(pos, 1ll << 56); s_fl[warp] = (wbad ? GSC_PF_BAD : 0) | (wzero ? GSC_PF_ZERO : 0); }
    __syncthreads();
    if (tid == 0) {
        long long f0 = 0, f1 = 0, ng = 0, ps = 0;
        int fl = GSC_PF_ZERO;
        for (int w = 0; w < GSC_PT / 32; ++w) {
            gsc_pfd_compose(f0, f1, s_f0[w], s_f1[w]);
            ng = min(ng + s_neg[w], 1ll << 57); ps = min(ps + s_pos[w], 1ll << 57);
            fl = (fl & s_fl[w] & GSC_PF_ZERO) | ((fl | s_fl[w]) & GSC_PF_BAD);
        }
        GscWin o;
        o.f0 = f0; o.f1 = f1; o.neg = ng; o.pos = ps; o.e0 = valid ? e0 : -1; o.flags = fl; o.pad = 0;
        win[blockIdx.x] = o;
    }
}

// ---- the chain: exact sequential sum of a[0..n) through the window summaries (one warp) ---------------------------
// out[0] = the exact sum; stats[0] += windows added up element by element
__global__ void __launch_bounds__(32) k_plan_chain(const double *__restrict__ a, long long n, const GscWin *__restrict__ win, long long nw,
                                                   double *__restrict__ out, unsigned long long *__restrict__ stats) {
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ __align__(16) double s_buf[GSC_PW];
    const int lane = threadIdx.x;
    double run = 0.0;
    unsigned long long n_exact = 0;
    for (long long wb = 0; wb < nw; wb += 32) {
        const long long wl = wb + lane;
        const int cntw = (int)min(32ll, nw - wb);
        GscWin w;
        w.f0 = 0; w.f1 = 0; w.neg = 0; w.pos = 0; w.e0 = -2; w.flags = GSC_PF_BAD;
        if (wl < nw) w = win[wl];
        int start = 0;
        while (start < cntw) {
            const long long rb = __double_as_longlong(run);
            const int ea = (int)((rb >> 52) & 0x7ff);
            const long long S0 = (rb & 0xfffffffffffffll) | (1ll << 52);
            const bool normal = rb > 0 && ea >= 64 && ea <= 2040;
            const bool in = lane >= start && lane < cntw;
            int firstbad = start;
            {
                // windows of zeros leave the sum alone whatever it is; the others need a summary made for this binade
                const bool zero = in && (w.flags & GSC_PF_ZERO);
                const bool cand = in && !zero && normal && !(w.flags & GSC_PF_BAD) && w.e0 == ea;
                long long s0 = cand ? w.f0 : 0, s1 = cand ? w.f1 : 0;
#pragma unroll
                for (int of = 1; of < 32; of <<= 1) {
                    const long long p0 = __shfl_up_sync(FULL, s0, of), p1 = __shfl_up_sync(FULL, s1, of);
                    if (lane >= of) { long long q0 = p0, q1 = p1; gsc_pfd_compose(q0, q1, s0, s1); s0 = q0; s1 = q1; }
                }
                long long x0 = __shfl_up_sync(FULL, s0, 1), x1 = __shfl_up_sync(FULL, s1, 1);
                if (lane == 0) { x0 = 0; x1 = 0; }
                const long long Sen = S0 + ((S0 & 1) ? x1 : x0);
                const bool ok = !in || zero || (cand && Sen - w.neg >= (1ll << 52) && Sen + w.pos < (1ll << 53));
                const unsigned badm = __ballot_sync(FULL, !ok);
                firstbad = badm ? (__ffs(badm) - 1) : cntw;
                if (normal) {
                    const long long e0 = __shfl_sync(FULL, s0, 31), e1 = __shfl_sync(FULL, s1, 31);
                    const long long Sfb = __shfl_sync(FULL, Sen, firstbad & 31);
                    run = gsc_mkd(ea, (firstbad < cntw) ? Sfb : S0 + ((S0 & 1) ? e1 : e0));
                }
                // (not normal: only zero windows were accepted, the sum is unchanged)
            }
            if (firstbad < cntw) {
                const long long w0 = (wb + firstbad) * GSC_PW;
                const int cnt = (int)min((long long)GSC_PW, n - w0);
                for (int e = lane; e < GSC_PW; e += 32) s_buf[e] = (e < cnt) ? a[w0 + e] : 0.0;
                __syncwarp();
                double r = run;
                if (lane == 0) {
                    for (int e = 0; e < cnt; e += 2) {
                        const double2 x = *reinterpret_cast<const double2 *>(&s_buf[e]);
                        r = r + x.x;
                        if (e + 1 < cnt) r = r + x.y;
                    }
                    ++n_exact;
                }
                run = __shfl_sync(FULL, r, 0);
                __syncwarp();
                start = firstbad + 1;
            } else {
                start = cntw;
            }
        }
    }
    if (lane == 0) { out[0] = run; if (stats) stats[0] += n_exact; }
}

// avg = sqrt(A / (S * C))   (enc:1386);   per = T / frameCount (enc:1400)
__global__ void k_plan_avg(const double *__restrict__ A, double count, double *__restrict__ avg) { *avg = sqrt(*A / count); }
__global__ void k_plan_per(const double *__restrict__ T, double frame_count, double *__restrict__ per) { *per = *T / frame_count; }

// ---- boundaries ------------------------------------------------------------------------------------------------
// Candidates from the approximate prefix, frame after frame from frame `k0` (whose start is known exactly): one warp.
// P(i) = wpre[w] + (tree) sum of the window's elements up to i; the frame that starts at b ends at the first
// i > b, i % block == 0, with P(i) - P(b) >= per.
__global__ void __launch_bounds__(32) k_plan_candidates(const double *__restrict__ t, long long S, const double *__restrict__ wpre,
                                                        const double *__restrict__ perp, int block, long long *__restrict__ starts,
                                                        int k0, int max_frames, int *__restrict__ n_frames) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    const double per = *perp;
    const long long nw = (S + GSC_PW - 1) / GSC_PW;
    auto prefix_at = [&](long long i) -> double {     // approximate sum of t[0..i], all lanes get it
        const long long w = i / GSC_PW, w0 = w * GSC_PW;
        double s = 0.0;
        for (long long j = w0 + lane; j <= i; j += 32) s += t[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
        return wpre[w] + s;
    };
    int k = k0;
    long long b = starts[k0];
    for (;;) {
        if (k + 1 >= max_frames) break;
        const double pb = (k == 0) ? 0.0 : prefix_at(b);      // frame 0's sum starts with t[0], the others behind their first sample
        const double target = pb + per;
        // first window whose END prefix reaches the target
        long long lo = b / GSC_PW, hi = nw;           // answer in [lo, nw]
        while (lo < hi) { const long long mid = (lo + hi) >> 1; if (wpre[mid + 1] >= target) hi = mid; else lo = mid + 1; }
        if (lo >= nw) break;                          // the rest of the file is the last frame
        // inside (and, for the block alignment, just behind) that window: first aligned i > b with P(i) >= target
        long long found = -1;
        for (long long w = lo; w < nw && found < 0 && w <= lo + 1; ++w) {
            const long long w0 = w * GSC_PW;
            double run = wpre[w];
            for (long long c0 = w0; c0 < min(S, w0 + GSC_PW) && found < 0; c0 += 32) {
                const long long i = c0 + lane;
                double v = (i < S) ? t[i] : 0.0, inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += u; }
                const double pi = run + inc;
                const bool hit = i < S && i > b && (i % block) == 0 && pi >= target;
                const unsigned m = __ballot_sync(FULL, hit);
                if (m) found = c0 + (__ffs(m) - 1);
                run += __shfl_sync(FULL, inc, 31);
            }
        }
        if (found < 0) break;
        ++k;
        b = found;
        if (lane == 0) starts[k] = b;
    }
    if (lane == 0) *n_frames = k + 1;
}

// One warp per frame runs the reference's loop (enc:1402-1423) exactly from the frame's start: cur = 0; for i = start + 1 ..:
// cur += t[i]; stop at the first i % block == 0 with cur >= per.  (The reference resets cur at a boundary sample AFTER
// adding that sample's term, so a frame's sum starts with the sample behind its first one; frame 0 starts with t[0].)
// exact_next[k] = that i, or S if the file ends first.
__global__ void __launch_bounds__(32) k_plan_verify(const double *__restrict__ t, long long S, const double *__restrict__ perp, int block,
                                                    const long long *__restrict__ starts, int n_frames, long long *__restrict__ exact_next) {
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ __align__(16) double s_buf[256];
    const int k = blockIdx.x, lane = threadIdx.x;
    if (k >= n_frames) return;
    const double per = *perp;
    const long long b = starts[k];
    const long long first = (k == 0) ? 0 : b + 1;
    // where the candidate says the frame ends (+ slack): the loop normally stops there; it may run on to the end of the file
    double cur = 0.0;
    long long found = S;
    for (long long c0 = first; c0 < S && found == S; c0 += 256) {
        const int cnt = (int)min(256ll, S - c0);
        for (int e = lane; e < 256; e += 32) s_buf[e] = (e < cnt) ? t[c0 + e] : 0.0;
        __syncwarp();
        long long f = S;
        double r = cur;
        if (lane == 0) {
            for (int e = 0; e < cnt; ++e) {
                r = r + s_buf[e];
                const long long i = c0 + e;
                if ((i % block) == 0 && r >= per) { f = i; break; }
            }
        }
        cur = __shfl_sync(FULL, r, 0);
        found = __shfl_sync(FULL, f, 0);
        __syncwarp();
    }
    if (lane == 0) exact_next[k] = found;
}
