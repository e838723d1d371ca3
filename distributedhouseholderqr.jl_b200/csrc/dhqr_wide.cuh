// dhqr_wide.cuh — the 128-column panel chain: CholeskyQR2 + Householder reconstruction on a whole outer panel.
//
// Replaces, for full aligned panels, four cooperative 32-column panel launches + three inner block updates (S:127-135 and
// S:198-213 for the columns inside the panel) by stream-ordered kernels with THREE grid-wide reductions per 128 columns and no
// spinning CTAs (tests/widepanel_model.py restates the stages in numpy; same reflectors as the reference's recurrences):
//
//   pack      P -> vpk (packed working copy, stays the V operand of the trailing GEMMs)
//   gram      G1 = P'P                (k_gemm_vta<128> with no trailing columns + k_wreduce)
//   chol128   R1 = chol(G1); Z1 = blocked inverse operand of R1                                one CTA
//   rmul      vpk <- vpk R1^{-1}  (= Q1): row-local blocked triangular solve on the fp64 tensor pipe, 64-row chunks
//   gram      G2 = Q1'Q1;  gram2_finish: guard |G2 - I|, R2 = I + U, Z2 = I - U to first order (U = striu(E) + diag(E)/2,
//             E = G2 - I) when max|E| <= 1e-9, which is every panel that is not nearly rank deficient; else chol128 again
//   trimm     Rt = R2 R1
//   rmul      top two chunks <- Q1top R2^{-1} (= Wt)
//   hr128     signed LU of Wt (Householder reconstruction), top block of the output (V, R, alpha), Rr     one CTA
//   trimm_z   Z23 = blocked inverse operand of Rr R2
//   rmul      rows below the top block: vpk <- Q1 (Rr R2)^{-1} (= V), also written to the caller's matrix
//
// "Blocked inverse operand" Z of an upper triangular R (32-column blocks): Z_bb = inv(R_bb), Z_ab = -R_ab inv(R_bb) (a < b), so
// that X = P R^{-1} is X_b = P_b Z_bb + sum_{a<b} X_a Z_ab: GEMM-shaped, row local, and only 32 x 32 blocks are inverted.
//
// The guards (positive finite Cholesky pivots, a conditioning estimate of R1, ||Q1'Q1 - I|| <= 1/4) are evaluated on the device;
// a refused panel records its index in WideCtl::fail_step, every later kernel that would write the caller's matrix returns at
// once, and the driver redoes the factorisation from that panel with the 32-column chain (dhqr_api.cu: qr_blocked).
#pragma once
#include "dhqr_kernels.cuh"

namespace dhqr {


// rmul operand layout of an upper-triangular 128 x 128 matrix X (B operand of vpk <- vpk X): per 32-column block nbk only
// the rows k < 32 (nbk + 1) are kept,  XL[xl_off(nbk) + (n % 32) * xl_ld(nbk) + k];  every leading dimension is == 4 mod 16
// so the DMMA B fragments load without bank conflicts; one bulk copy brings the whole operand into shared memory.
__host__ __device__ __forceinline__ constexpr int xl_ld(int nbk) { return 32 * (nbk + 1) + 4; }
__host__ __device__ __forceinline__ constexpr int xl_off(int nbk) { return nbk == 0 ? 0 : (nbk == 1 ? 1152 : (nbk == 2 ? 3328 : 6528)); }
constexpr int XL_ELEMS = 10752;    // 32 * (36 + 68 + 100 + 132)
constexpr double FIRST_ORDER_MAX = 1e-9;   // second pass: max |Q1'Q1 - I| accepted (chol(I + E) to first order in E)

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
// chol128: R = chol(G) (upper) and its blocked inverse operand Z; one CTA of 512 threads.  G: [j * 128 + i] (k_wreduce).
//   Guards: positive finite pivots and the conditioning estimate
//       est^2 = sum_b ||diag(||p_j||) Z_bb||_F^2 + sum_{a<b} ||Z_ab||_F^2 <= kappa_max^2
//   (inverting diagonal blocks explicitly costs ~1e-17 x est in ||QR - A|| / ||A||, invariant under column scaling:
//   tests/test_widepanel_model.py).
//   The trailing matrix lives in registers: thread (warp w, lane l) holds rows w + 16 a, columns l + 32 b.  Step j: the warp
//   that owns row j scales it (one rsqrt) and publishes it through shared memory, one barrier, and every thread updates its
//   8 x 4 block (the symmetric update needs row j only).  The published rows are R.
// ------------------------------------------------------------------------------------------------
constexpr int WT = 4 * 32 * LDD;   // scratch: four inverted diagonal blocks
constexpr size_t SMEM_WIDE1 = ((size_t)WP * WLD + WT + 8 * WP) * 8 + 64;

__global__ void __launch_bounds__(512, 1) k_chol128(const double* __restrict__ G, double* __restrict__ Rp,
                                                    double* __restrict__ ZL, WideCtl* ctl, int step, double* vflag,
                                                    double kappa_max, long long* stamps) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* A = reinterpret_cast<double*>(smem_raw);   // [128][WLD] row-major: R
    double* T = A + WP * WLD;                            // WT: the four inverted diagonal blocks
    double* rinv = T + WT;                               // 128
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
                bad |= !(fabs(g[a][b]) < 1e300);
                if (i == k) dn[i] = sqrt(g[a][b]);
            }
        if (bad) *sbad = 1;
    }
    __syncthreads();
    if (stamps && tid == 0) stamps[0] = clock64() - t0;
    if (!*sbad) {
        double dcur = 1.0, rcur = 1.0;                 // pivot of the row this warp publishes next and its rsqrt
#pragma unroll
        for (int aj = 0; aj < 8; ++aj) {
            const int bj = aj >> 1;
            if (warp == 0) {                           // first row of the block of 16: nothing to overlap the rsqrt with
                dcur = __shfl_sync(0xffffffffu, g[aj][bj], (aj & 1) * 16);
                rcur = rsqrt_nb(dcur);
            }
#pragma unroll 1
            for (int t = 0; t < 16; ++t) {
                const int j = 16 * aj + t, lj = (aj & 1) * 16 + t;
                double* rb = rowbuf + (j & 1) * WP;
                if (warp == t) {                       // this warp holds row j in g[aj][*]
                    const double d = dcur, ri = rcur;
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
                        if (a == aj && warp == t + 1) {
                            // this warp publishes row j + 1 next: its pivot is final now; the rsqrt (the long dependent chain of a
                            // step) runs while the warp updates its other rows instead of after the next barrier
                            dcur = __shfl_sync(0xffffffffu, g[aj][bj], lj + 1);
                            rcur = rsqrt_nb(dcur);
                        }
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
        if (warp < 4) triu_inv32_warp(A + (32 * warp) * WLD + 32 * warp, WLD, rinv + 32 * warp, T + warp * 32 * LDD, lane);
        __syncthreads();
        if (stamps && tid == 0) stamps[2] = clock64() - t0;
        double acc = 0.0;
        for (int e = tid; e < 4096; e += 512) {          // Z_bb = inv(R_bb)
            const int b = e >> 10, c = (e >> 5) & 31, i = e & 31;
            const double v = T[b * 32 * LDD + i * LDD + c];
            ZL[xl_off(b) + c * xl_ld(b) + 32 * b + i] = v;
            const double sv = dn[32 * b + i] * v;
            acc += sv * sv;
        }
        for (int e = tid; e < 6144; e += 512) {          // Z_ab = -R_ab inv(R_bb), a < b
            const int pr = e >> 10, c = (e >> 5) & 31, i = e & 31;
            const int a = pr < 3 ? 0 : (pr < 5 ? 1 : 2), b = pr < 3 ? pr + 1 : (pr < 5 ? pr - 1 : 3);
            const double* r = A + (32 * a + i) * WLD + 32 * b;
            const double* d = T + b * 32 * LDD + c;
            double s0 = 0.0, s1 = 0.0;
            int p = 0;
            for (; p + 1 <= c; p += 2) {
                s0 += r[p] * d[p * LDD];
                s1 += r[p + 1] * d[(p + 1) * LDD];
            }
            if (p <= c) s0 += r[p] * d[p * LDD];
            const double v = -(s0 + s1);
            ZL[xl_off(b) + c * xl_ld(b) + 32 * a + i] = v;
            acc += v * v;
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
    if (*sbad && tid == 0) {
        ctl->status = 1;
        atomicMin(&ctl->fail_step, step);
        if (vflag) *vflag = 1.0;
    }
    if (stamps && tid == 0) stamps[3] = clock64() - t0;
}

// ------------------------------------------------------------------------------------------------
// gram2_finish: the split-K reduction of the second Gram matrix (as k_wreduce, fixed order) fused with the second Cholesky
// pass in its first-order form.  E = G2 - I is the loss of orthogonality of the first pass; with U = striu(E) + diag(E)/2,
// chol(I + E) = I + U + O(E^2) and its inverse is I - U + O(E^2).  max|E| <= 1e-9 bounds the neglected terms by
// 128 * 1e-18, far below rounding.  E is O(eps kappa^2): a panel that passed the conditioning guard of the first pass
// (kappa <~ 1e3) sits orders of magnitude below the bound; anything larger REFUSES the panel (the driver redoes it with the
// 32-column chain), so the chain carries no second full Cholesky.
// One element per thread (grid 64 x 256).  Outputs: Ws (G2), Rp = R2 plain, ZL = its blocked inverse operand.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gram2_finish(const double* __restrict__ Wp, int64_t pstride, int nsplit, double* __restrict__ Ws,
                                                      double* __restrict__ Rp, double* __restrict__ ZL, WideCtl* ctl, int step,
                                                      double* vflag) {
    if (wide_gate_closed(ctl, step) || ctl->status) return;
    int e;
    double g;
    if (gridDim.x == 64) {                               // one element per thread
        e = blockIdx.x * 256 + threadIdx.x;
        if (e >= WP * WP) return;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int p = 0;
        for (; p + 4 <= nsplit; p += 4) {
            s0 += Wp[(int64_t)p * pstride + e];
            s1 += Wp[(int64_t)(p + 1) * pstride + e];
            s2 += Wp[(int64_t)(p + 2) * pstride + e];
            s3 += Wp[(int64_t)(p + 3) * pstride + e];
        }
        for (; p < nsplit; ++p) s0 += Wp[(int64_t)p * pstride + e];
        g = (s0 + s1) + (s2 + s3);
    } else {                                             // grid 256: four lanes per element (the order of k_wreduce4)
        const int t = blockIdx.x * 256 + threadIdx.x, q = t & 3;
        e = t >> 2;
        double s = 0.0;
        int p = q;
        for (; p + 12 < nsplit; p += 16) {
            const double v0 = Wp[(int64_t)p * pstride + e], v1 = Wp[(int64_t)(p + 4) * pstride + e];
            const double v2 = Wp[(int64_t)(p + 8) * pstride + e], v3 = Wp[(int64_t)(p + 12) * pstride + e];
            s += v0; s += v1; s += v2; s += v3;
        }
        for (; p < nsplit; p += 4) s += Wp[(int64_t)p * pstride + e];
        const double s1 = __shfl_xor_sync(0xffffffffu, s, 1);
        const double a2 = (q & 1) ? (s1 + s) : (s + s1);
        g = a2 + __shfl_xor_sync(0xffffffffu, a2, 2);
        if (q != 0) return;
    }
    Ws[e] = g;
    const int i = e & (WP - 1), j = e >> 7;              // row, column
    const double E = g - (i == j ? 1.0 : 0.0);
    if (!(fabs(E) <= FIRST_ORDER_MAX)) {              // also catches NaN / Inf
        ctl->status = 1;
        atomicMin(&ctl->fail_step, step);
        if (vflag) *vflag = 1.0;
    }
    const double u = i == j ? 0.5 * E : E;
    Rp[e] = i < j ? u : (i == j ? 1.0 + u : 0.0);
    const int nbk = j >> 5;
    if (i < 32 * (nbk + 1)) ZL[xl_off(nbk) + (j & 31) * xl_ld(nbk) + i] = i < j ? -u : (i == j ? 1.0 - u : 0.0);
}

// ------------------------------------------------------------------------------------------------
// trimm128: C = A B for upper-triangular 128 x 128 operands (plain column-major, zeros below the diagonal);
// grid = the 10 upper 32 x 32 blocks; every block loads all its operand blocks in one go (one global round trip).
// ------------------------------------------------------------------------------------------------
constexpr size_t SMEM_TRIMM = (size_t)11 * 32 * 33 * 8;

__device__ __forceinline__ void trimm_block_id(int bid, int& ib, int& jb) {
    ib = 0;
    int r = bid;
    while (r >= 4 - ib) { r -= 4 - ib; ++ib; }
    jb = ib + r;
}

__global__ void __launch_bounds__(256) k_trimm128(const double* __restrict__ Am, const double* __restrict__ Bm,
                                                  double* __restrict__ Cp, const WideCtl* ctl, int step) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sA = reinterpret_cast<double*>(smem_raw);   // [4][32][33]
    double* sB = sA + 4 * 32 * 33;
    if (wide_gate_closed(ctl, step) || ctl->status) return;
    int ib, jb;
    trimm_block_id(blockIdx.x, ib, jb);
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
    for (int q = 0; q < 4; ++q) Cp[(size_t)(jb * 32 + jq + 8 * q) * WP + ib * 32 + i] = acc[q];
    // the blocks below the diagonal are zero in the plain layout; whoever owns the diagonal block of a block column clears them
    if (ib == jb)
        for (int pb = jb + 1; pb < 4; ++pb)
#pragma unroll
            for (int q = 0; q < 4; ++q) Cp[(size_t)(jb * 32 + jq + 8 * q) * WP + pb * 32 + i] = 0.0;
}

// trimm_z: Z = blocked inverse operand of C = A B (A, B upper triangular, plain), without forming C in memory: block (a, b)
// of the grid computes C_ab and C_bb, inverts C_bb with one warp, and writes Z_bb = inv(C_bb) (a == b) or Z_ab = -C_ab inv(C_bb).
__global__ void __launch_bounds__(256) k_trimm_z(const double* __restrict__ Am, const double* __restrict__ Bm, double* __restrict__ ZL,
                                                 const WideCtl* ctl, int step) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sA = reinterpret_cast<double*>(smem_raw);   // [4][32][33]: A(a, a..b)
    double* sB = sA + 4 * 32 * 33;                       // [4][32][33]: B(a..b, b)
    double* sA2 = sB + 4 * 32 * 33;                      // A(b, b)
    double* sC = sA2 + 32 * 33;                          // C_ab, then C_bb
    double* sD = sC + 32 * 33;                           // inv(C_bb), [32][LDD]
    if (wide_gate_closed(ctl, step) || ctl->status) return;
    int ib, jb;
    trimm_block_id(blockIdx.x, ib, jb);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, i = lane, jq = warp;
    for (int pb = ib; pb <= jb; ++pb)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = jq + 8 * q;
            sA[((pb - ib) * 32 + i) * 33 + c] = Am[(size_t)(pb * 32 + c) * WP + ib * 32 + i];
            sB[((pb - ib) * 32 + i) * 33 + c] = Bm[(size_t)(jb * 32 + c) * WP + pb * 32 + i];
        }
#pragma unroll
    for (int q = 0; q < 4; ++q) sA2[i * 33 + jq + 8 * q] = Am[(size_t)(jb * 32 + jq + 8 * q) * WP + jb * 32 + i];
    __syncthreads();
    double cab[4] = {0.0, 0.0, 0.0, 0.0}, cbb[4] = {0.0, 0.0, 0.0, 0.0};
    for (int pb = 0; pb <= jb - ib; ++pb) {
        const double* a = sA + (pb * 32 + i) * 33;
        const double* bq = sB + pb * 32 * 33 + jq;
#pragma unroll 8
        for (int p = 0; p < 32; ++p) {
            const double av = a[p];
#pragma unroll
            for (int q = 0; q < 4; ++q) cab[q] += av * bq[p * 33 + 8 * q];
        }
    }
    {
        const double* a = sA2 + i * 33;
        const double* bq = sB + (jb - ib) * 32 * 33 + jq;
#pragma unroll 8
        for (int p = 0; p < 32; ++p) {
            const double av = a[p];
#pragma unroll
            for (int q = 0; q < 4; ++q) cbb[q] += av * bq[p * 33 + 8 * q];
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) sC[i * 33 + jq + 8 * q] = cbb[q];
    __syncthreads();
    if (warp == 0) triu_inv32_warp(sC, 33, nullptr, sD, lane);
    __syncthreads();
    if (ib == jb) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = jq + 8 * q;
            ZL[xl_off(jb) + c * xl_ld(jb) + 32 * jb + i] = sD[i * LDD + c];
        }
        return;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) sC[i * 33 + jq + 8 * q] = cab[q];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int c = jq + 8 * q;
        double s = 0.0;
        for (int p = 0; p <= c; ++p) s += sC[i * 33 + p] * sD[p * LDD + c];
        ZL[xl_off(jb) + c * xl_ld(jb) + 32 * ib + i] = -s;
    }
}

// ------------------------------------------------------------------------------------------------
// vpk_rmul:  chunks [q0, q0 + nq) of vpk  <-  chunk * R^{-1} through the blocked inverse operand Z of R (ZL layout), on the
// fp64 tensor pipe; optionally the result also goes to the caller's matrix (rows < mp of the panel at P).
//   The solve is row local: warp w owns rows 8w .. 8w+7 of the 64-row chunk and runs the four 32-column block steps
//   X_b = P_b Z_bb + sum_{a<b} X_a Z_ab by itself (the finished blocks overwrite the chunk in shared memory and are the A
//   operand of the later steps: __syncwarp only).  A CTA keeps Z in shared memory and walks over its chunks with two chunk
//   buffers: the bulk copy of the next chunk and the bulk store of the previous result overlap the DMMAs of the current one.
// ------------------------------------------------------------------------------------------------
struct RmulArgs {
    double* vpk;
    int q0, nq;
    const double* ZL;
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
            for (int o = 0; o < XL_ELEMS; o += 3584) bulk_g2s(sX + o, a.ZL + o, 3584 * 8, &bar[buf]);
        const double* src = a.vpk + (int64_t)qq * VPK_CHUNK;
        double* dst = sC0 + buf * VPK_CHUNK;
        for (int o = 0; o < VPK_CHUNK; o += VPK_CHUNK / 4) bulk_g2s(dst + o, src + o, VPK_CHUNK * 2, &bar[buf]);
    };
    if (q < qend && tid == 0) load(q, 0, true);
    for (int it = 0; q < qend; q += qstep, ++it) {
        const int buf = it & 1;
        double* sC = sC0 + buf * VPK_CHUNK;
        if (tid == 0 && q + qstep < qend) {
            bulk_wait_read0();                               // the store issued from the other buffer has finished reading it
            load(q + qstep, buf ^ 1, false);
        }
        mbar_wait(&bar[buf], (it >> 1) & 1);
        const double* pa = sC + (lane & 3) * LD1 + 8 * warp + (lane >> 2);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int ld = xl_ld(b), k4 = 8 * (b + 1);
            const double* pb = sX + xl_off(b) + (lane >> 2) * ld + (lane & 3);
            double acc[4][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = 0.0;
#pragma unroll 4
            for (int kk = 0; kk < k4; ++kk) {
                const double a0 = pa[kk * 4 * LD1];
                double bf[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = pb[j * 8 * ld + kk * 4];
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[j][0], acc[j][1], a0, bf[j]);
            }
            __syncwarp();                                     // every lane has read P_b of the warp's rows
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int row = 8 * warp + (lane >> 2), col = 32 * b + 8 * j + 2 * (lane & 3);
                sC[col * LD1 + row] = acc[j][0];
                sC[(col + 1) * LD1 + row] = acc[j][1];
            }
            __syncwarp();                                     // X_b is the A operand of the next block steps
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
//   The rows below the top block are V = Q Rr^{-1}, Rr = diag(sqrt(U)) (I + diag(-S/U) striu(W)), left in Rrp (plain).
//   MTp = U_lu' D'^{-1} (lower triangular, plain) with U_lu the upper factor of the LU of E - Q S and D' = diag(v_jj): the
//   right-hand side from which k_trecon gets the compact-WY factor, T' = V1^{-1} MT.
//   Like k_chol128 the matrix lives in registers (rows w + 16 a, columns l + 32 b per thread); a step publishes column j and
//   row j through double-buffered shared memory: one barrier per step.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1) k_hr128(double* __restrict__ vpk, const double* __restrict__ Rt, double* __restrict__ P,
                                                  int64_t ldp, double* __restrict__ alpha, double* __restrict__ Rrp,
                                                  double* __restrict__ MTp, const WideCtl* ctl, int step, long long* stamps) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* Wt = reinterpret_cast<double*>(smem_raw);   // [128][WLD] row-major
    double* T = Wt + WP * WLD;
    double* Sg = T + WT;
    double* Ud = Sg + WP;
    double* rsq = Ud + WP;
    double* colbuf = rsq + WP;                           // [2][128]
    double* rowbuf = colbuf + 2 * WP;                    // [2][128]
    double* fb = rowbuf + 2 * WP;                        // [2]
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
    double fcur = 0.0, scur = 0.0, ucur = 1.0;         // S_j / U_jj, S_j, U_jj of the pivot this thread publishes next
#pragma unroll
    for (int aj = 0; aj < 8; ++aj) {
        const int bj = aj >> 1;
        if (warp == 0 && lane == (aj & 1) * 16) {          // first pivot of the block of 16
            const double pv = w[aj][bj];
            scur = pv > 0.0 ? -1.0 : 1.0;
            ucur = 1.0 + fabs(pv);
            fcur = scur / ucur;
        }
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
                if (lane == lj) { fb[j & 1] = fcur; Sg[j] = scur; Ud[j] = ucur; }
            }
            __syncthreads();
            const double f = fb[j & 1];
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
                    if (a == aj && warp == t + 1 && lane == lj + 1) {
                        // the next pivot is final: its division runs while this thread updates its other rows
                        const double pv = w[aj][bj];
                        scur = pv > 0.0 ? -1.0 : 1.0;
                        ucur = 1.0 + fabs(pv);
                        fcur = scur / ucur;
                    }
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
        const double wij = Wt[i * WLD + j];
        double v, rr, mt;
        if (i > j) {
            v = wij * rsq[j];
            rr = 0.0;
            mt = Wt[j * WLD + i] * Sg[i] * Sg[j] * rsq[j];       // frozen row j, column i: U_lu(j, i) / D'_j
        } else if (i == j) {
            v = mt = -Sg[j] * (Ud[j] * rsq[j]);
            rr = Ud[i] * rsq[i];
        } else {
            v = Sg[i] * Rt[e];
            rr = (-Sg[i] / Ud[i]) * wij * (Ud[i] * rsq[i]);
            mt = 0.0;
        }
        P[(int64_t)j * ldp + i] = v;
        vpk[vpk_index(i, j)] = i >= j ? v : 0.0;
        Rrp[e] = rr;
        MTp[e] = mt;
        if (i == j) alpha[j] = Sg[j] * Rt[e];
    }
    if (stamps && tid == 0) stamps[6] = clock64() - t0;
}

// ------------------------------------------------------------------------------------------------
// trecon: the compact-WY factor of the panel from the reconstruction itself.  I - V T V' restricted to the top block reads
// V1 T V1' = E - Q S = L_lu U_lu, hence T' = V1^{-1} (U_lu' D'^{-1}) = V1^{-1} MT (V1 = top 128 x 128 block of V, lower
// triangular; MT from k_hr128): a triangular solve with 128 right-hand sides instead of the 128 x 128 Gram matrix V'V over all
// rows (2 rows 128^2 flop inside k_gemm_vta) + k_tinv on the chain.  Block column jb of T' per CTA:
//   T'(ib, jb) = inv(V1_ii) (MT(ib, jb) - sum_{pb = jb}^{ib-1} V1(ib, pb) T'(pb, jb)),   ib = jb .. 3.
// Output: Linv[j * 128 + i] = T'(i, j) (what k_ymake reads).  Verified against (I + stril(V'V))^{-1} to 2e-16 in numpy.
// ------------------------------------------------------------------------------------------------
constexpr size_t SMEM_TRECON = (size_t)(10 + 4 + 1 + 4) * 32 * 33 * 8;

__global__ void __launch_bounds__(256) k_trecon(const double* __restrict__ vpk, const double* __restrict__ MTp, double* __restrict__ Linv,
                                                const WideCtl* ctl, int step) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sV = reinterpret_cast<double*>(smem_raw);   // [10][32][33]: V1(ib, pb), pb <= ib, block index ib (ib + 1) / 2 + pb
    double* sT = sV + 10 * 32 * 33;                      // [4][32][33]: T'(ib, jb)
    double* sA = sT + 4 * 32 * 33;                       // [32][33]: right-hand side of the current step
    double* sD = sA + 32 * 33;                           // [4][32][LDD]: inv(V1_ii')  (upper; element (c, r) = inv(V1_ii)(r, c))
    if (wide_gate_closed(ctl, step) || ctl->status) return;
    const int jb = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, i = lane, jq = warp;
    for (int ib = jb; ib < 4; ++ib)
        for (int pb = jb; pb <= ib; ++pb)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int cc = jq + 8 * q;
                sV[((ib * (ib + 1) / 2 + pb) * 32 + i) * 33 + cc] = vpk[vpk_index(32 * ib + i, 32 * pb + cc)];
            }
    __syncthreads();
    if (warp < 4 - jb) {                                  // inv(V1_ii) for ii = jb + warp: invert the transpose (upper) by one warp
        const int ii = jb + warp;
        double* U = sT + warp * 32 * 33;                  // scratch: sT is not live yet
        const double* Vd = sV + ((ii * (ii + 1) / 2 + ii) * 32) * 33;
        for (int r = 0; r < 32; ++r) U[lane * 33 + r] = r >= lane ? Vd[r * 33 + lane] : 0.0;     // U(c = lane, r) = V1_ii(r, c)
        __syncwarp();
        triu_inv32_warp(U, 33, nullptr, sD + ii * 32 * LDD, lane);
    }
    __syncthreads();
    for (int ib = jb; ib < 4; ++ib) {
        double acc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] = MTp[(size_t)(32 * jb + jq + 8 * q) * WP + 32 * ib + i];
        for (int pb = jb; pb < ib; ++pb) {
            const double* a = sV + ((ib * (ib + 1) / 2 + pb) * 32 + i) * 33;
            const double* t = sT + pb * 32 * 33 + jq;
#pragma unroll 8
            for (int p = 0; p < 32; ++p) {
                const double av = a[p];
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] -= av * t[p * 33 + 8 * q];
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) sA[i * 33 + jq + 8 * q] = acc[q];
        __syncthreads();
        const double* d = sD + ib * 32 * LDD;             // inv(V1_ii)(i, p) = d[p * LDD + i], p <= i
        double out[4] = {0.0, 0.0, 0.0, 0.0};
        for (int p = 0; p <= i; ++p) {
            const double dv = d[p * LDD + i];
#pragma unroll
            for (int q = 0; q < 4; ++q) out[q] += dv * sA[p * 33 + jq + 8 * q];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int cc = jq + 8 * q;
            sT[ib * 32 * 33 + i * 33 + cc] = out[q];
            Linv[(size_t)(32 * jb + cc) * WP + 32 * ib + i] = out[q];
        }
        __syncthreads();
    }
    // blocks above the diagonal of T' are zero (k_ymake only reads the lower triangle, tools read the whole matrix)
    for (int ib = 0; ib < jb; ++ib)
#pragma unroll
        for (int q = 0; q < 4; ++q) Linv[(size_t)(32 * jb + jq + 8 * q) * WP + 32 * ib + i] = 0.0;
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
