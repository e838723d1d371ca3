// dhqr_wide.cuh — the 128-column panel chain: CholeskyQR2 + Householder reconstruction on a whole outer panel.
//
// Replaces, for full aligned panels, four cooperative 32-column panel launches + three inner block updates (S:127-135 and
// S:198-213 for the columns inside the panel) by stream-ordered kernels with THREE grid-wide reductions per 128 columns and no
// spinning CTAs (tests/widepanel_model.py restates the stages in numpy; same reflectors as the reference's recurrences):
//
//   pack      P -> vpk (packed working copy, stays the V operand of the trailing GEMMs)
//   gram      G1 = P'P                (k_gemm_vta<128> with no trailing columns + k_wreduce)
//   chol128   R1 = chol(G1), X1 = R1^{-1}                       one CTA
//   rmul      vpk <- vpk X1  (= Q1)                              DMMA, one 64-row chunk per CTA iteration
//   gram      G2 = Q1'Q1
//   chol128   guard |G2 - I|, R2 = chol(G2), X2 = R2^{-1}
//   trimm     Rt = R2 R1
//   rmul      top two chunks <- Q1top X2 (= Wt)
//   hr128     signed LU of Wt (Householder reconstruction), top block of the output (V, R, alpha), Y3 = Rr^{-1}
//   trimm     X3 = X2 Y3
//   rmul      rows below the top block: vpk <- Q1 X3 (= V), also written to the caller's matrix
//
// The guards (positive finite Cholesky pivots, ||Q1'Q1 - I|| <= 1/4) are evaluated on the device; a refused panel records its
// index in WideCtl::fail_step, every later kernel that would write the caller's matrix returns at once, and the driver redoes
// the factorisation from that panel with the 32-column chain (dhqr_api.cu: qr_blocked).
#pragma once
#include "dhqr_kernels.cuh"

namespace dhqr {

constexpr int WP = 128;            // wide panel width
constexpr int WLD = WP + 1;        // leading dimension of the row-major 128 x 128 work matrices in shared memory

// rmul operand layout of an upper-triangular 128 x 128 matrix X (B operand of vpk <- vpk X): per 32-column block nbk only
// the rows k < 32 (nbk + 1) are kept,  XL[xl_off(nbk) + (n % 32) * xl_ld(nbk) + k];  every leading dimension is == 4 mod 16
// so the DMMA B fragments load without bank conflicts; one bulk copy brings the whole operand into shared memory.
__host__ __device__ __forceinline__ constexpr int xl_ld(int nbk) { return 32 * (nbk + 1) + 4; }
__host__ __device__ __forceinline__ constexpr int xl_off(int nbk) { return nbk == 0 ? 0 : (nbk == 1 ? 1152 : (nbk == 2 ? 3328 : 6528)); }
constexpr int XL_ELEMS = 10752;    // 32 * (36 + 68 + 100 + 132)

__device__ __forceinline__ double rsqrt_nb(double d) {
    // rsqrt(double) without the library's slow-path branch: MUFU seed + one cubic step (the fast path of rsqrt())
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
    const double t = y0 * y0;
    const double e = fma(d, -t, 1.0);
    const double pq = fma(e, 0.375, 0.5);
    const double q = y0 * e;
    return fma(pq, q, y0);
}

__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// In-place inverse of the upper triangle of A (row-major, leading dimension WLD, 128 x 128) by recursive doubling:
// 8 x 8 diagonal blocks by back substitution (dinv = 1 / diag), then X12 = -X11 (R12 X22) for block sizes 8 .. 64.
// The strict lower triangle is neither read nor written.  T: scratch of 4096 doubles.  All threads of the CTA call it.
// (Four partial sums per dot product: the loops are latency bound, one warp instruction in flight per sum.)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void triu_inv128(double* A, const double* dinv, double* T, int tid, int nthreads) {
    if (tid < WP) {
        const int d = tid >> 3, c = tid & 7;
        const double* Rb = A + (d * 8) * WLD + d * 8;
        double x[8];
#pragma unroll
        for (int i = 7; i >= 0; --i) {
            double s = i == c ? 1.0 : 0.0;
#pragma unroll
            for (int p = i + 1; p < 8; ++p)
                if (p <= c) s -= Rb[i * WLD + p] * x[p];
            x[i] = i <= c ? s * dinv[d * 8 + i] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) T[d * 64 + i * 8 + c] = x[i];
    }
    __syncthreads();
    for (int e = tid; e < 16 * 64; e += nthreads) {
        const int d = e >> 6, i = (e >> 3) & 7, c = e & 7;
        if (i <= c) A[(d * 8 + i) * WLD + d * 8 + c] = T[e];
    }
    __syncthreads();
#pragma unroll 1
    for (int lb = 3; lb < 7; ++lb) {
        const int bs = 1 << lb, per = bs * bs, tot = (WP / (2 * bs)) * per;
        // tmp = R12 X22 : lanes run over the row i (the trip count depends on the column j only)
        for (int e = tid; e < tot; e += nthreads) {
            const int pi = e >> (2 * lb), r = e & (per - 1), i = r & (bs - 1), j = r >> lb, o = pi * 2 * bs;
            const double* r12 = A + (o + i) * WLD + o + bs;
            const double* x22 = A + (o + bs) * WLD + o + bs + j;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int p = 0;
            for (; p + 3 <= j; p += 4) {
                s0 += r12[p] * x22[p * WLD];
                s1 += r12[p + 1] * x22[(p + 1) * WLD];
                s2 += r12[p + 2] * x22[(p + 2) * WLD];
                s3 += r12[p + 3] * x22[(p + 3) * WLD];
            }
            for (; p <= j; ++p) s0 += r12[p] * x22[p * WLD];
            T[pi * per + i * bs + j] = (s0 + s1) + (s2 + s3);
        }
        __syncthreads();
        // X12 = -X11 tmp : lanes run over the column j (the trip count depends on the row i only)
        for (int e = tid; e < tot; e += nthreads) {
            const int pi = e >> (2 * lb), r = e & (per - 1), j = r & (bs - 1), i = r >> lb, o = pi * 2 * bs;
            const double* x11 = A + (o + i) * WLD + o;
            const double* tm = T + pi * per + j;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int p = i;
            for (; p + 3 < bs; p += 4) {
                s0 += x11[p] * tm[p * bs];
                s1 += x11[p + 1] * tm[(p + 1) * bs];
                s2 += x11[p + 2] * tm[(p + 2) * bs];
                s3 += x11[p + 3] * tm[(p + 3) * bs];
            }
            for (; p < bs; ++p) s0 += x11[p] * tm[p * bs];
            A[(o + i) * WLD + o + bs + j] = -((s0 + s1) + (s2 + s3));
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// chol128: R = chol(G) (upper), X = R^{-1}; one CTA of 512 threads.  G: [j * 128 + i] as k_wreduce leaves it.
//   second == 0: guards = positive finite pivots and ||D X||_F <= kappa_max, D = diag(||p_j||) (the explicit inverse costs
//                ~1e-17 x that number in ||QR - A|| / ||A||, scale invariant: tests/test_widepanel_model.py)
//   second != 0: guard max |G - I| <= 1 / (4 * 128) first (the first pass left Q1 close enough to orthonormal for the second
//                pass to finish the job; refuses NaN / Inf as well).
//   The trailing matrix lives in registers: thread (warp w, lane l) holds rows w + 16 a, columns l + 32 b.  Step j: the warp
//   that owns row j scales it (one rsqrt) and publishes it through shared memory, one barrier, and every thread updates its
//   8 x 4 block (the symmetric update needs row j only).  The published rows are R; X by recursive doubling afterwards.
//   Outputs: Rp, Xp plain column-major upper triangular (zeros below), XL the rmul operand layout of X.
// ------------------------------------------------------------------------------------------------
constexpr size_t SMEM_WIDE1 = ((size_t)WP * WLD + 4096 + 8 * WP) * 8 + 64;

__global__ void __launch_bounds__(512, 1) k_chol128(const double* __restrict__ G, int second, double* __restrict__ Rp,
                                                    double* __restrict__ Xp, double* __restrict__ XL, WideCtl* ctl, int step,
                                                    double* vflag, double kappa_max, long long* stamps) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* A = reinterpret_cast<double*>(smem_raw);   // [128][WLD] row-major: R, then X
    double* T = A + WP * WLD;                            // 4096
    double* rinv = T + 4096;                             // 128
    double* dn = rinv + WP;                              // 128: ||p_j|| = sqrt(G_jj)
    double* rowbuf = dn + WP;                            // [2][128]
    double* red = rowbuf + 2 * WP;                       // 16 (+ spare)
    int* sbad = reinterpret_cast<int*>(rinv + 8 * WP);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (wide_gate_closed(ctl, step) || ctl->status) return;
    const long long t0 = clock64();
    if (tid == 0) *sbad = 0;
    double g[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) g[a][b] = G[(warp + 16 * a) * WP + lane + 32 * b];   // G is symmetric: (k, i) for (i, k)
    __syncthreads();
    {
        int bad = 0;
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int i = warp + 16 * a, k = lane + 32 * b;
                if (second) bad |= !(fabs(g[a][b] - (i == k ? 1.0 : 0.0)) <= 0.25 / WP);
                else bad |= !(fabs(g[a][b]) < 1e300);
                if (i == k) dn[i] = sqrt(g[a][b]);
            }
        if (bad) *sbad = 1;
    }
    __syncthreads();
    if (stamps && tid == 0) stamps[0] = clock64() - t0;
    if (!*sbad) {
#pragma unroll
        for (int aj = 0; aj < 8; ++aj) {
            const int bj = aj >> 1;
#pragma unroll 1
            for (int t = 0; t < 16; ++t) {
                const int j = 16 * aj + t, lj = (aj & 1) * 16 + t;
                double* rb = rowbuf + (j & 1) * WP;
                if (warp == t) {                       // this warp holds row j in g[aj][*]
                    const double d = __shfl_sync(0xffffffffu, g[aj][bj], lj);
                    const double ri = rsqrt_nb(d);
                    if (lane == 0) {
                        if (!(d > 0.0) || !(d < 1e300)) *sbad = 1;
                        rinv[j] = ri;
                    }
#pragma unroll
                    for (int b = bj; b < 4; ++b) {
                        const int k = lane + 32 * b;
                        const double r = k == j ? d * ri : (k > j ? g[aj][b] * ri : 0.0);
                        rb[k] = r;
                        A[j * WLD + k] = r;
                    }
                }
                __syncthreads();
                double rc[4];
#pragma unroll
                for (int b = bj; b < 4; ++b) rc[b] = rb[lane + 32 * b];
#pragma unroll
                for (int a = aj; a < 8; ++a)
                    if (a > aj || warp > t) {
                        const double rr = rb[warp + 16 * a];
#pragma unroll
                        for (int b = bj; b < 4; ++b) g[a][b] -= rr * rc[b];
                    }
            }
        }
    }
    __syncthreads();
    if (stamps && tid == 0) stamps[1] = clock64() - t0;
    if (!*sbad) {
        for (int e = tid; e < WP * WP; e += 512) {
            const int i = e & (WP - 1), j = e >> 7;
            Rp[e] = i <= j ? A[i * WLD + j] : 0.0;
        }
        __syncthreads();
        triu_inv128(A, rinv, T, tid, 512);
        if (stamps && tid == 0) stamps[2] = clock64() - t0;
        if (!second) {                                   // ||D X||_F
            double acc = 0.0;
            for (int e = tid; e < WP * WP; e += 512) {
                const int k = e >> 7, n = e & (WP - 1);
                if (k <= n) {
                    const double v = dn[k] * A[k * WLD + n];
                    acc += v * v;
                }
            }
            acc = warp_sum(acc);
            if (lane == 0) red[warp] = acc;
            __syncthreads();
            if (tid == 0) {
                double tsum = 0.0;
                for (int w = 0; w < 16; ++w) tsum += red[w];
                if (!(tsum <= kappa_max * kappa_max)) *sbad = 1;
            }
            __syncthreads();
        }
    }
    if (*sbad) {
        if (tid == 0) {
            ctl->status = 1;
            atomicMin(&ctl->fail_step, step);
            if (vflag) *vflag = 1.0;
        }
        return;
    }
    if (Xp)
        for (int e = tid; e < WP * WP; e += 512) {
            const int i = e & (WP - 1), j = e >> 7;
            Xp[e] = i <= j ? A[i * WLD + j] : 0.0;
        }
#pragma unroll
    for (int nbk = 0; nbk < 4; ++nbk) {
        const int ld = xl_ld(nbk), kk = 32 * (nbk + 1);
        for (int e = tid; e < 32 * kk; e += 512) {
            const int k = e % kk, nin = e / kk, n = 32 * nbk + nin;
            XL[xl_off(nbk) + nin * ld + k] = k <= n ? A[k * WLD + n] : 0.0;
        }
    }
    if (stamps && tid == 0) stamps[3] = clock64() - t0;
}

// ------------------------------------------------------------------------------------------------
// trimm128: C = A B for upper-triangular 128 x 128 operands (plain column-major, zeros below the diagonal);
// grid = the 10 upper 32 x 32 blocks; every block loads all its operand blocks in one go (one global round trip).
// Outputs: Cp plain and / or CL in the rmul operand layout (either may be null).
// ------------------------------------------------------------------------------------------------
constexpr size_t SMEM_TRIMM = (size_t)8 * 32 * 33 * 8;

__global__ void __launch_bounds__(256) k_trimm128(const double* __restrict__ Am, const double* __restrict__ Bm,
                                                  double* __restrict__ Cp, double* __restrict__ CL, const WideCtl* ctl, int step) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sA = reinterpret_cast<double*>(smem_raw);   // [4][32][33]
    double* sB = sA + 4 * 32 * 33;
    if (wide_gate_closed(ctl, step) || ctl->status) return;
    int ib = 0, r = blockIdx.x;
    while (r >= 4 - ib) { r -= 4 - ib; ++ib; }
    const int jb = ib + r;
    const int tid = threadIdx.x, i = tid & 31, jq = tid >> 5;   // thread: row i, columns jq, jq + 8, jq + 16, jq + 24
    for (int pb = ib; pb <= jb; ++pb)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = jq + 8 * q;
            sA[((pb - ib) * 32 + i) * 33 + c] = Am[(size_t)(pb * 32 + c) * WP + ib * 32 + i];   // A(ib, pb): (i, c)
            sB[((pb - ib) * 32 + i) * 33 + c] = Bm[(size_t)(jb * 32 + c) * WP + pb * 32 + i];   // B(pb, jb): (i, c)
        }
    __syncthreads();
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int pb = 0; pb <= jb - ib; ++pb) {
        const double* a = sA + (pb * 32 + i) * 33;
        const double* bq = sB + pb * 32 * 33 + jq;
#pragma unroll 8
        for (int p = 0; p < 32; ++p) {
            const double av = a[p];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] += av * bq[p * 33 + 8 * q];
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int row = ib * 32 + i, cin = jq + 8 * q, col = jb * 32 + cin;
        if (Cp) Cp[(size_t)col * WP + row] = acc[q];
        if (CL) CL[xl_off(jb) + cin * xl_ld(jb) + row] = acc[q];
    }
    // the blocks below the diagonal are zero in the plain layout; whoever owns the diagonal block of a block column clears them
    if (Cp && ib == jb)
        for (int pb = jb + 1; pb < 4; ++pb)
#pragma unroll
            for (int q = 0; q < 4; ++q) Cp[(size_t)(jb * 32 + jq + 8 * q) * WP + pb * 32 + i] = 0.0;
}

// ------------------------------------------------------------------------------------------------
// vpk_rmul:  chunks [q0, q0 + nq) of vpk  <-  chunk * X   (X upper triangular 128 x 128 in the XL layout), on the fp64
// tensor pipe; optionally the result also goes to the caller's matrix (rows < mp of the panel at P).
//   A CTA keeps X in shared memory and walks over its chunks with two chunk buffers: the bulk copy of the next chunk and the
//   bulk store of the previous result overlap the DMMAs of the current one.  8 warps x (16 rows x two 32-column blocks paired
//   (0,3) / (1,2) so that every warp runs the same number of k steps of the triangular product); the result overwrites the
//   chunk in shared memory and goes back with one bulk store.
// ------------------------------------------------------------------------------------------------
struct RmulArgs {
    double* vpk;
    int q0, nq;
    const double* XL;
    double* P;          // null: packed output only
    int64_t ldp, mp;
    const WideCtl* ctl;
    int step;
};
constexpr size_t SMEM_RMUL = ((size_t)XL_ELEMS + 2 * VPK_CHUNK) * 8 + 64;

__global__ void __launch_bounds__(256, 1) k_vpk_rmul(RmulArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sX = reinterpret_cast<double*>(smem_raw);
    double* sC0 = sX + XL_ELEMS;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sC0 + 2 * VPK_CHUNK);   // [2]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (wide_gate_closed(a.ctl, a.step) || a.ctl->status) return;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    const int qend = a.q0 + a.nq, qstep = gridDim.x;
    int q = a.q0 + blockIdx.x;
    auto load = [&](int qq, int buf, bool withX) {          // thread 0 only
        const uint32_t cb = VPK_CHUNK * 8;
        mbar_arrive_expect_tx(&bar[buf], cb + (withX ? XL_ELEMS * 8 : 0));
        if (withX)
            for (int o = 0; o < XL_ELEMS; o += 3584) bulk_g2s(sX + o, a.XL + o, 3584 * 8, &bar[buf]);
        const double* src = a.vpk + (int64_t)qq * VPK_CHUNK;
        double* dst = sC0 + buf * VPK_CHUNK;
        for (int o = 0; o < VPK_CHUNK; o += VPK_CHUNK / 4) bulk_g2s(dst + o, src + o, VPK_CHUNK * 2, &bar[buf]);
    };
    if (q < qend && tid == 0) load(q, 0, true);
    const int rb = warp & 3, pr = warp >> 2;
    for (int it = 0; q < qend; q += qstep, ++it) {
        const int buf = it & 1;
        double* sC = sC0 + buf * VPK_CHUNK;
        if (tid == 0 && q + qstep < qend) {
            bulk_wait_read0();                               // the store issued from the other buffer has finished reading it
            load(q + qstep, buf ^ 1, false);
        }
        mbar_wait(&bar[buf], (it >> 1) & 1);
        double acc[2][2][4][2];
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[b][i][j][0] = acc[b][i][j][1] = 0.0;
        const double* pa = sC + (lane & 3) * LD1 + 16 * rb + (lane >> 2);
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int nbk = b == 0 ? pr : 3 - pr;
            const int ld = xl_ld(nbk), k4 = 8 * (nbk + 1);
            const double* pb = sX + xl_off(nbk) + (lane >> 2) * ld + (lane & 3);
#pragma unroll 4
            for (int kk = 0; kk < k4; ++kk) {
                const double a0 = pa[kk * 4 * LD1], a1 = pa[kk * 4 * LD1 + 8];
                double bf[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = pb[j * 8 * ld + kk * 4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dmma(acc[b][0][j][0], acc[b][0][j][1], a0, bf[j]);
                    dmma(acc[b][1][j][0], acc[b][1][j][1], a1, bf[j]);
                }
            }
        }
        __syncthreads();                                      // every warp has read its A fragments
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int nbk = b == 0 ? pr : 3 - pr;
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int row = 16 * rb + 8 * i + (lane >> 2), col = 32 * nbk + 8 * j + 2 * (lane & 3);
                    sC[col * LD1 + row] = acc[b][i][j][0];
                    sC[(col + 1) * LD1 + row] = acc[b][i][j][1];
                }
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            double* dst = a.vpk + (int64_t)q * VPK_CHUNK;
            for (int o = 0; o < VPK_CHUNK; o += VPK_CHUNK / 4) bulk_s2g(dst + o, sC + o, VPK_CHUNK * 2);
            bulk_commit();
        }
        if (a.P) {
            for (int c = warp; c < WP; c += 8)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int64_t row = (int64_t)q * KC1 + lane + 32 * h;
                    if (row < a.mp) a.P[(int64_t)c * a.ldp + row] = sC[c * LD1 + lane + 32 * h];
                }
            __syncthreads();                                  // the generic reads of this buffer end before it is reloaded
        }
    }
    if (tid == 0) bulk_wait0();
}

// ------------------------------------------------------------------------------------------------
// hr128: Householder reconstruction of the top block; one CTA of 512 threads.
//   Wt = first 128 rows of vpk (= rows of the orthonormal factor Q2).  Signed LU, row j frozen at step j:
//   S_j = -sign(w_jj), U_jj = 1 + |w_jj|, W(i,k) += (S_j / U_jj) W(i,j) W(j,k).  In the reference's storage (S:127-135):
//   v_ij = W_ij^(j) / sqrt(U_jj) (i > j), v_jj = -S_j sqrt(U_jj), alpha_j = S_j Rt_jj, R_ij = S_i Rt_ij (i < j).
//   The rows below the top block are V = Q Rr^{-1}, Rr = diag(sqrt(U)) (I + diag(-S/U) striu(W)); this kernel leaves
//   Y3 = Rr^{-1} (plain) for k_trimm128 / k_vpk_rmul.
//   Like k_chol128 the matrix lives in registers (rows w + 16 a, columns l + 32 b per thread); a step publishes column j and
//   row j through double-buffered shared memory: one barrier per step.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1) k_hr128(double* __restrict__ vpk, const double* __restrict__ Rt, double* __restrict__ P,
                                                  int64_t ldp, double* __restrict__ alpha, double* __restrict__ Y3,
                                                  const WideCtl* ctl, int step, long long* stamps) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* Wt = reinterpret_cast<double*>(smem_raw);   // [128][WLD] row-major
    double* T = Wt + WP * WLD;
    double* Sg = T + 4096;
    double* Ud = Sg + WP;
    double* rsq = Ud + WP;
    double* colbuf = rsq + WP;                           // [2][128]
    double* rowbuf = colbuf + 2 * WP;                    // [2][128]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (wide_gate_closed(ctl, step) || ctl->status) return;
    const long long t0 = clock64();
    for (int e = tid; e < WP * WP; e += 512) {
        const int i = e & (WP - 1), k = e >> 7;
        Wt[i * WLD + k] = vpk[vpk_index(i, k)];
    }
    __syncthreads();
    double w[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) w[a][b] = Wt[(warp + 16 * a) * WLD + lane + 32 * b];
    if (stamps && tid == 0) stamps[4] = clock64() - t0;
#pragma unroll
    for (int aj = 0; aj < 8; ++aj) {
        const int bj = aj >> 1;
#pragma unroll 1
        for (int t = 0; t < 16; ++t) {
            const int j = 16 * aj + t, lj = (aj & 1) * 16 + t;
            double* cb = colbuf + (j & 1) * WP;
            double* rb = rowbuf + (j & 1) * WP;
            if (lane == lj) {
#pragma unroll
                for (int a = aj; a < 8; ++a) cb[warp + 16 * a] = w[a][bj];
            }
            if (warp == t) {
#pragma unroll
                for (int b = bj; b < 4; ++b) rb[lane + 32 * b] = w[aj][b];
            }
            __syncthreads();
            const double pv = rb[j];
            const double sgn = pv > 0.0 ? -1.0 : 1.0;
            const double u = 1.0 + fabs(pv);
            if (tid == 0) { Sg[j] = sgn; Ud[j] = u; }
            const double f = sgn / u;
            double rc[4];
#pragma unroll
            for (int b = bj; b < 4; ++b) rc[b] = rb[lane + 32 * b];
#pragma unroll
            for (int a = aj; a < 8; ++a)
                if (a > aj || warp > t) {
                    const double li = f * cb[warp + 16 * a];
#pragma unroll
                    for (int b = bj; b < 4; ++b)
                        if (b > bj || lane > lj) w[a][b] += li * rc[b];
                }
        }
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) Wt[(warp + 16 * a) * WLD + lane + 32 * b] = w[a][b];
    if (tid < WP) rsq[tid] = 1.0 / sqrt(Ud[tid]);
    __syncthreads();
    if (stamps && tid == 0) stamps[5] = clock64() - t0;
    for (int e = tid; e < WP * WP; e += 512) {
        const int i = e & (WP - 1), j = e >> 7;
        double v;
        if (i > j) v = Wt[i * WLD + j] * rsq[j];
        else if (i == j) v = -Sg[j] * (Ud[j] * rsq[j]);
        else v = Sg[i] * Rt[e];
        P[(int64_t)j * ldp + i] = v;
        vpk[vpk_index(i, j)] = i >= j ? v : 0.0;
        if (i == j) alpha[j] = Sg[j] * Rt[e];
    }
    __syncthreads();
    for (int e = tid; e < WP * WP; e += 512) {           // Rr over the upper triangle, in place
        const int i = e >> 7, k = e & (WP - 1);
        if (k >= i) {
            const double sq = Ud[i] * rsq[i];
            Wt[i * WLD + k] = k == i ? sq : (-Sg[i] / Ud[i]) * Wt[i * WLD + k] * sq;
        }
    }
    __syncthreads();
    triu_inv128(Wt, rsq, T, tid, 512);
    for (int e = tid; e < WP * WP; e += 512) {
        const int i = e & (WP - 1), j = e >> 7;
        Y3[e] = i <= j ? Wt[i * WLD + j] : 0.0;
    }
    if (stamps && tid == 0) stamps[6] = clock64() - t0;
}

// start of a wide panel: clear the guards of the previous one and the validity flag that travels with the V buffer
__global__ void k_wide_begin(WideCtl* ctl, double* vflag) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        ctl->status = 0;
        *vflag = 0.0;
    }
}
__global__ void k_wide_reset(WideCtl* ctl) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        ctl->fail_step = W_NOFAIL;
        ctl->status = 0;
    }
}
// after a V buffer arrived from another rank: take over the owner's verdict on the panel
__global__ void k_wide_note(WideCtl* ctl, const double* vflag, int step) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && *vflag != 0.0) atomicMin(&ctl->fail_step, step);
}

}  // namespace dhqr
