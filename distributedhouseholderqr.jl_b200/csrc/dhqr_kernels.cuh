// dhqr_kernels.cuh — hand-written sm_100a kernels for the blocked Householder QR hot path.
//
// Reference semantics (S:n = /root/reference/src/DistributedHouseholderQR.jl:n):
//   column step      S:127-135   s=|x|, alpha=-sign(x1)s, f=1/sqrt(s(s+|x1|)), v=f(x-alpha e1), |v|^2=2
//   trailing update  S:198-213   a <- a - v (v'a)      (partialdot S:42-49, hotloop! S:156-160)
//   Q'b sweep        S:232-242
//   back-substitute  S:256-282
// The kernels here compute the same reflectors, but blocked: nb reflectors are aggregated into
//   Q_panel' = I - V T' V',  T^{-1} = I + striu(V'V)        (beta == 1 because |v|^2 == 2)
// so that the trailing update is two dense fp64 GEMMs on the tensor pipe (DMMA; tcgen05 has no
// f64 kind), fed by TMA bulk copies (cp.async.bulk -> UBLKCP) through an mbarrier ring.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dhqr {

constexpr int KC = 32;        // K-chunk: rows per stage in gemm_vta, V columns per stage in gemm_cvy
constexpr int LDK = KC + 4;   // padded leading dim of [column][k] smem tiles: 36 doubles (288 B), 36 % 16 == 4
constexpr int IB = 32;        // inner (cooperative) panel width
constexpr int PANEL_THREADS = 512;
#ifndef DHQR_PANEL_VARIANT
#define DHQR_PANEL_VARIANT 4   // bit 2: triangular solves of the panel fast path on the fp64 tensor pipe (0: row-by-row
#endif                         // substitution on the vector pipe); A/B results in profiles/r01_panel_variants.txt.
constexpr int PANEL_VARIANT = DHQR_PANEL_VARIANT;
// Fast-path guard on the first Cholesky factor: min / max of its diagonal.  Row-by-row substitution is backward stable for
// any factor the other guards accept (1e-5); the blocked solves invert 8x8 diagonal blocks explicitly, which costs
// ~5e-18 x spread in ||QR - A|| / ||A|| (tests/test_fastpath_model.py), so they only take panels with a spread below 250.
constexpr double FAST_SPREAD_MIN = (PANEL_VARIANT & 4) ? 4e-3 : 1e-5;

// Control words of the speculative 128-column panel chain (dhqr_wide.cuh).  fail_step = index of the first outer panel whose
// guards refused the fast factorisation (W_NOFAIL: none); every kernel that writes the caller's matrix carries a `gate` and
// returns at once when a panel with an index below its gate has failed, so that the driver can redo the factorisation from
// that panel on an untouched trailing matrix.
constexpr int W_NOFAIL = 0x7fffffff;
struct WideCtl {
    int fail_step;
    int status;      // guards of the panel in flight: 0 = fine
    int pad[2];
};
__device__ __forceinline__ bool wide_gate_closed(const WideCtl* ctl, int gate) {
    return ctl && *reinterpret_cast<const volatile int*>(&ctl->fail_step) < gate;
}

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier, TMA bulk copy, fp64 tensor-core MMA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar), ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP).
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// D(8x8) += A(8x4, row) * B(4x8, col), fp64 tensor pipe (SASS: DMMA.8x8x4).
//   a : A[lane>>2][lane&3]      b : B[lane&3][lane>>2]      c0,c1 : C[lane>>2][2*(lane&3) + {0,1}]
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Consumer-side release of the PREVIOUS stage, called right after the wait for the current one.
// Releasing stage s at the end of its own chunk is not safe: ptxas hoists the arrive above the last
// DMMAs, i.e. directly behind the last LDS of the stage, and the TMA producer (async proxy) can then
// overwrite the buffer while that read is still in flight (observed: 8x32 blocks of C wrong once per
// ~5e6 CTAs).  One iteration later every DMMA of the previous chunk has issued, hence every LDS it
// depends on has returned.  Costs nothing: the producer refills the stage during the current chunk.
__device__ __forceinline__ void release_prev_stage(uint64_t* empty, int it, int stages, int lane) {
    if (it > 0) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[(it - 1) % stages]);
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// Packed operand layouts (handle-owned buffers, written by k_panel / k_pack / k_ymake)
//   vpk : Householder block V of the current outer panel, 64-row chunks, each chunk stored exactly as
//         the padded shared-memory tile the GEMMs want:   vpk[q][col][LD1],  q = window_row / 64,
//         col in [0,128), LD1 = 68 (68 % 16 == 4 -> conflict-free DMMA fragment loads).
//         => gemm_vta stages a whole V chunk, and gemm_cvy a 64 x 32 slice, with ONE TMA bulk copy.
//   ypk : Y = -T'W,  ypk[n_tile][k_chunk][64 cols][LDK]  -> one bulk copy per gemm_cvy stage.
// (Issuing one 256 B bulk copy per column and stage made gemm_vta TMA-issue bound at 36 % of peak.)
// ------------------------------------------------------------------------------------------------
constexpr int KC1 = 64;                    // rows per gemm_vta stage
constexpr int LD1 = KC1 + 4;               // 68
constexpr int VPK_COLS = 128;
constexpr int VPK_CHUNK = VPK_COLS * LD1;  // doubles per 64-row chunk
constexpr int YT = 64;                     // ypk column-tile width (== gemm_cvy BN)
constexpr int YCOLS = 32;                  // columns per k_ymake / k_mid32 CTA

__device__ __host__ __forceinline__ int64_t vpk_index(int64_t wrow, int col) {
    return ((wrow >> 6) * VPK_COLS + col) * LD1 + (wrow & 63);
}

// ------------------------------------------------------------------------------------------------
// gemm_vta:  Wext(NBP x next) = V' * [V | A]      "TN", reduction over the long row dimension
//   V   : NBP packed columns [voff, voff+NBP) of vpk; one bulk copy per 64-row chunk.
//   A   : trailing columns in user storage, `rows` valid rows, 512 B bulk copy per column and chunk
//         (generic loads for a ragged tail or unaligned storage), spread over NPW producer warps.
//   grid: (tiles over ext columns, splits over row chunks); each CTA writes one partial tile to
//         Wp[split]; k_wreduce sums the partials in a fixed order (deterministic).
//   CTA : WM*WN consumer warps (32x32 warp tiles of 8x8x4 DMMAs) + NPW TMA producer warps, 2 stages.
// ------------------------------------------------------------------------------------------------
struct GemmVtaArgs {
    const double* vpk;  // packed V, window row 0
    int voff;           // first packed column of this V block
    int nv;             // leading ext columns taken from V itself (Gram block S = V'V); == NBP
    const double* A;    // window row 0, first trailing column
    int64_t lda;
    int64_t rows;       // valid rows of A in the window
    int na;             // trailing columns
    int nchunks;        // ceil(rows / KC1)
    int a_aligned;      // 1: every A column start is 16B aligned (bulk copies legal)
    double* Wp;         // partials: [split][next_pad][NBP]
    int64_t pstride;    // elements between consecutive partials
};

template <int NBP, int BN, int WM, int WN, int NPW>
__global__ void __launch_bounds__((WM * WN + NPW) * 32, 1) k_gemm_vta(GemmVtaArgs a) {
    constexpr int NCW = WM * WN;
    constexpr int STAGES = 2;
    constexpr int WTM = NBP / WM, WTN = BN / WN;
    constexpr int MI = WTM / 8, NJ = WTN / 8;
    constexpr int CPW = BN / NPW;   // B columns per producer warp
    static_assert(WTM % 8 == 0 && WTN % 8 == 0 && BN % NPW == 0, "tile");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sV = reinterpret_cast<double*>(smem_raw);   // [STAGES][NBP][LD1]
    double* sB = sV + STAGES * NBP * LD1;                // [STAGES][BN][LD1]
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * BN * LD1);
    uint64_t* empty = full + STAGES;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int next = a.nv + a.na;
    const int col0 = blockIdx.x * BN;
    const int ncols_tile = min(BN, next - col0);
    const int cps = (a.nchunks + gridDim.y - 1) / gridDim.y;
    const int ch0 = blockIdx.y * cps;
    const int nit = max(min(ch0 + cps, a.nchunks) - ch0, 0);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], NPW);
            mbar_init(&empty[s], NCW);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp >= NCW) {
        // ===== TMA producer warps =====
        const int pw = warp - NCW;
        const int cbeg = pw * CPW, cend = min(cbeg + CPW, ncols_tile);
        for (int it = 0; it < nit; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            const int64_t q = ch0 + it;
            const int64_t krow = q * KC1;
            const int64_t left = a.rows - krow;
            const int nvalid = left >= KC1 ? KC1 : (int)left;        // valid A rows in this chunk (>= 1)
            const int nbulk = a.a_aligned ? (nvalid & ~1) : 0;        // rows moved by TMA per A column
            double* dV = sV + (size_t)s * NBP * LD1;
            double* dB = sB + (size_t)s * BN * LD1;
            const double* vchunk = a.vpk + q * VPK_CHUNK;
            // generic-proxy fill of what TMA cannot move (ragged tail / unaligned user storage)
            for (int c = cbeg + lane; c < cend; c += 32) {
                const int jg = col0 + c;
                if (jg >= a.nv && nbulk < KC1) {
                    const double* src = a.A + (int64_t)(jg - a.nv) * a.lda + krow;
                    double* dst = dB + c * LD1;
                    for (int r = nbulk; r < KC1; ++r) dst[r] = r < nvalid ? src[r] : 0.0;
                }
            }
            const int nvc = max(min(a.nv - col0, cend) - cbeg, 0);    // columns of this warp that come from V
            const int nac = max(cend - cbeg, 0) - nvc;                // ... and from A
            uint32_t bytes = (uint32_t)nvc * KC1 * 8 + (uint32_t)nac * nbulk * 8;
            if (pw == 0) bytes += NBP * LD1 * 8;
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_expect_tx(&full[s], bytes);
                if (pw == 0) bulk_g2s(dV, vchunk + (int64_t)a.voff * LD1, NBP * LD1 * 8, &full[s]);
            }
            __syncwarp();
            for (int c = cbeg + lane; c < cend; c += 32) {
                const int jg = col0 + c;
                if (jg < a.nv) {
                    bulk_g2s(dB + c * LD1, vchunk + (int64_t)(a.voff + jg) * LD1, KC1 * 8, &full[s]);
                } else if (nbulk > 0) {
                    bulk_g2s(dB + c * LD1, a.A + (int64_t)(jg - a.nv) * a.lda + krow, nbulk * 8, &full[s]);
                }
            }
        }
        return;
    }

    // ===== DMMA consumer warps =====
    const int wm = warp / WN, wn = warp % WN;
    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int frag = (lane >> 2) * LD1 + (lane & 3);
    for (int it = 0; it < nit; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full[s], ph);
        release_prev_stage(empty, it, STAGES, lane);
        const double* v = sV + (size_t)s * NBP * LD1 + wm * WTM * LD1 + frag;
        const double* b = sB + (size_t)s * BN * LD1 + wn * WTN * LD1 + frag;
#pragma unroll 4
        for (int kk = 0; kk < KC1 / 4; ++kk) {
            double af[MI], bf[NJ];
#pragma unroll
            for (int i = 0; i < MI; ++i) af[i] = v[i * 8 * LD1 + kk * 4];
#pragma unroll
            for (int j = 0; j < NJ; ++j) bf[j] = b[j * 8 * LD1 + kk * 4];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    double* out = a.Wp + (int64_t)blockIdx.y * a.pstride;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int row = wm * WTM + i * 8 + (lane >> 2);
            const int col = col0 + wn * WTN + j * 8 + (lane & 3) * 2;
            if (col < next) out[(int64_t)col * NBP + row] = acc[i][j][0];
            if (col + 1 < next) out[(int64_t)(col + 1) * NBP + row] = acc[i][j][1];
        }
}

// ------------------------------------------------------------------------------------------------
// wreduce:  Ws[e] = sum_p Wp[p][e]   (fixed order -> deterministic), e over next*NBP elements
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wreduce(const double* __restrict__ Wp, int64_t pstride, int nsplit, int64_t nelem,
                                                 double* __restrict__ Ws) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += (int64_t)gridDim.x * blockDim.x) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int p = 0;
        for (; p + 4 <= nsplit; p += 4) {
            s0 += Wp[(int64_t)p * pstride + e];
            s1 += Wp[(int64_t)(p + 1) * pstride + e];
            s2 += Wp[(int64_t)(p + 2) * pstride + e];
            s3 += Wp[(int64_t)(p + 3) * pstride + e];
        }
        for (; p < nsplit; ++p) s0 += Wp[(int64_t)p * pstride + e];
        Ws[e] = (s0 + s1) + (s2 + s3);
    }
}

// wreduce4: the same sum for a SMALL block (a 128 x 128 Gram matrix) where k_wreduce is a chain of nsplit dependent-latency
// loads per thread: four lanes per element, lane q sums the partials p = q, q+4, ... (ascending), then ((q0+q1)+(q2+q3)).
// Fixed order -> deterministic (but not the order of k_wreduce).
__global__ void __launch_bounds__(256) k_wreduce4(const double* __restrict__ Wp, int64_t pstride, int nsplit, int64_t nelem,
                                                  double* __restrict__ Ws) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = t >> 2;
    const int q = (int)(t & 3);
    double s = 0.0;
    if (e < nelem) {
        int p = q;
        for (; p + 12 < nsplit; p += 16) {
            const double v0 = Wp[(int64_t)p * pstride + e], v1 = Wp[(int64_t)(p + 4) * pstride + e];
            const double v2 = Wp[(int64_t)(p + 8) * pstride + e], v3 = Wp[(int64_t)(p + 12) * pstride + e];
            s += v0; s += v1; s += v2; s += v3;
        }
        for (; p < nsplit; p += 4) s += Wp[(int64_t)p * pstride + e];
    }
    const double s1 = __shfl_xor_sync(0xffffffffu, s, 1);
    const double a = (q & 1) ? (s1 + s) : (s + s1);          // lanes 0,1 hold q0 + q1; lanes 2,3 hold q2 + q3
    const double a2 = __shfl_xor_sync(0xffffffffu, a, 2);
    if (q == 0 && e < nelem) Ws[e] = a + a2;
}

// ------------------------------------------------------------------------------------------------
// gram_sym:  partial Gram matrices of a packed 128-column panel,  G_s = sum over the CTA's 64-row chunks of Vc' Vc.
//   k_gemm_vta with the panel as both operands stages every chunk twice (once as V, once as the ext columns of each of its two
//   column tiles) and computes all 16 32x32 blocks; here a chunk is staged ONCE (one 68 KB bulk copy, 3-stage ring) and only the
//   10 blocks on or above the diagonal are computed, each by two warps (32 x 16 halves: 20 MMA warps = 5 per scheduler, balanced);
//   the off-diagonal blocks are written to both triangles, so the partials have the layout the consumers of k_gemm_vta's
//   partials expect ([split][column][128]).  grid = splits over the chunks; deterministic (fixed chunk order per CTA).
// ------------------------------------------------------------------------------------------------
constexpr int GS_STAGES = 3, GS_MMA_WARPS = 20;
constexpr size_t SMEM_GRAM_SYM = (size_t)GS_STAGES * VPK_CHUNK * 8 + 2 * GS_STAGES * 8;
struct GramSymArgs {
    const double* vpk;  // packed panel, window row 0 (rows padded with zeros to whole chunks)
    int nchunks;        // 64-row chunks to sum over
    double* Wp;         // partials: [split][128][128]
    int64_t pstride;
};
__global__ void __launch_bounds__((GS_MMA_WARPS + 1) * 32, 1) k_gram_sym(GramSymArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sV = reinterpret_cast<double*>(smem_raw);   // [GS_STAGES][128][LD1]
    uint64_t* full = reinterpret_cast<uint64_t*>(sV + (size_t)GS_STAGES * VPK_CHUNK);
    uint64_t* empty = full + GS_STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cps = (a.nchunks + gridDim.x - 1) / gridDim.x;
    const int ch0 = blockIdx.x * cps;
    const int nit = max(min(ch0 + cps, a.nchunks) - ch0, 0);
    if (tid == 0) {
        for (int s = 0; s < GS_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], GS_MMA_WARPS);
        }
        fence_mbar_init();
    }
    __syncthreads();
    if (warp == GS_MMA_WARPS) {
        if (lane == 0) {
            for (int it = 0; it < nit; ++it) {
                const int s = it % GS_STAGES;
                mbar_wait(&empty[s], ((it / GS_STAGES) & 1) ^ 1);
                mbar_arrive_expect_tx(&full[s], (uint32_t)(VPK_CHUNK * 8));
                bulk_g2s(sV + (size_t)s * VPK_CHUNK, a.vpk + (int64_t)(ch0 + it) * VPK_CHUNK, VPK_CHUNK * 8, &full[s]);
            }
        }
        return;
    }
    // warp -> (block row bi, block column bj >= bi, half h of the block's columns)
    const int blk = warp >> 1, h = warp & 1;
    int bi = 0, r = blk;
    while (r >= 4 - bi) { r -= 4 - bi; ++bi; }
    const int bj = bi + r;
    double acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int frag = (lane >> 2) * LD1 + (lane & 3);
    for (int it = 0; it < nit; ++it) {
        const int s = it % GS_STAGES;
        mbar_wait(&full[s], (it / GS_STAGES) & 1);
        release_prev_stage(empty, it, GS_STAGES, lane);
        const double* v = sV + (size_t)s * VPK_CHUNK + bi * 32 * LD1 + frag;
        const double* b = sV + (size_t)s * VPK_CHUNK + (bj * 32 + h * 16) * LD1 + frag;
#pragma unroll 4
        for (int kk = 0; kk < KC1 / 4; ++kk) {
            double af[4], bf[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = v[i * 8 * LD1 + kk * 4];
#pragma unroll
            for (int j = 0; j < 2; ++j) bf[j] = b[j * 8 * LD1 + kk * 4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    double* out = a.Wp + (int64_t)blockIdx.x * a.pstride;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int row = bi * 32 + i * 8 + (lane >> 2);
            const int col = bj * 32 + h * 16 + j * 8 + (lane & 3) * 2;
            out[(int64_t)col * VPK_COLS + row] = acc[i][j][0];
            out[(int64_t)(col + 1) * VPK_COLS + row] = acc[i][j][1];
            if (bi != bj) {                                            // the mirror image below the diagonal
                out[(int64_t)row * VPK_COLS + col] = acc[i][j][0];
                out[(int64_t)row * VPK_COLS + col + 1] = acc[i][j][1];
            }
        }
}

// ------------------------------------------------------------------------------------------------
// gemm_cvy:  C(rows x ncols) += V(rows x nbp) * Y(nbp x ncols)   on rows >= row_lo   ("NN", K = nbp)
//   Y already carries the minus sign and T' (ymake), so this is A_trail <- (I - V T' V') A_trail.
//   grid: (row tiles of 128, column tiles of 64); 2 CTAs per SM so one CTA's C-tile load/store
//   overlaps the other's MMA main loop.  CTA: 4 consumer warps (64x32 warp tiles) + 1 TMA warp.
//   Every stage is three bulk copies: two 64x32 slices of vpk and one 32x64 block of ypk.
// ------------------------------------------------------------------------------------------------
struct GemmCvyArgs {
    double* C;          // window row 0, first column
    int64_t ldc;
    int64_t rows;       // valid rows in the window
    int64_t row_lo;     // rows below this index (window-relative) are left untouched
    int ncols;
    const double* vpk;  // packed V, window row 0 (rows padded to a multiple of 128 with zeros)
    int voff;           // first packed column of this V block
    const double* ypk;  // packed Y: [n_tile][k_chunk][64][LDK]
    int nkq;            // k-chunks (of KC columns) to run
    int nkq_alloc;      // k-chunks per n_tile in ypk (tile stride)
    unsigned int* sm_ticket;   // [#SMs] ever-increasing per-SM counters (phase staggering), may be null
    int first_wave;     // CTAs with a linear id below this are in the first wave
    int stagger_cycles; // delay of the odd-ticket CTA of an SM in the first wave
    const WideCtl* ctl; // speculative panel chain: skip when a panel below `gate` was refused (may be null)
    int gate;
    int tiles_m, tiles_n;   // persistent variant: row tiles (of 128) x column tiles (of 64)
    int tiles_per_cta;      // persistent variant: consecutive tiles one CTA walks through before it retires
};

// DEFER (requires nkq == MI, i.e. the 128-wide update with 32-row warp tiles): the accumulators start at zero and the C tile is
// read in MI batches of one 8-row block each, batch i issued at the start of k-stage i and added when that stage's DMMAs are
// done, so the loads from HBM have a whole stage to arrive instead of stalling the warps before the first DMMA
// (profiles/r01_prof_cvy_ncu.txt: long_scoreboard 4.8 per issue, tensor pipe 79 % active with all loads up front).
template <int WM, int MINB, bool DEFER>
__global__ void __launch_bounds__((WM * 2 + 1) * 32, MINB) k_gemm_cvy(GemmCvyArgs a) {
    // WM = 2: 4 MMA warps with 64x32 warp tiles;  WM = 4: 8 MMA warps with 32x32 warp tiles (more warps per
    // scheduler to hide the C-tile loads/stores and the LDS latency)
    constexpr int BM = 128, BN = YT, WN = 2, NCW = WM * WN, STAGES = 2;
    constexpr int WTM = BM / WM, WTN = BN / WN;
    constexpr int MI = WTM / 8, NJ = WTN / 8;
    constexpr int VH = KC * LD1;   // doubles per 64-row x 32-col slice
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sV = reinterpret_cast<double*>(smem_raw);   // [STAGES][2][KC][LD1]
    double* sY = sV + STAGES * 2 * VH;                   // [STAGES][BN][LDK]
    uint64_t* full = reinterpret_cast<uint64_t*>(sY + STAGES * BN * LDK);
    uint64_t* empty = full + STAGES;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int nit = a.nkq;
    if (wide_gate_closed(a.ctl, a.gate)) return;

    // The two CTAs that share an SM start together and would stay phase-locked (both loading C, both
    // in the MMA loop, both storing): delay one of each first-wave pair by about half a tile so that
    // one CTA's C-tile traffic overlaps the other's tensor work for the rest of the kernel.
    if (a.sm_ticket && (int)(blockIdx.x + blockIdx.y * gridDim.x) < a.first_wave) {
        __shared__ unsigned int ticket;
        if (tid == 0) {
            unsigned int smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            ticket = atomicAdd(&a.sm_ticket[smid], 1u);
        }
        __syncthreads();
        if (ticket & 1u) {
            const long long t0 = clock64();
            while (clock64() - t0 < a.stagger_cycles) __nanosleep(200);
        }
    }

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NCW);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == NCW) {
        // ===== TMA producer warp =====
        if (lane == 0) {
            const double* v0 = a.vpk + (int64_t)(2 * blockIdx.x) * VPK_CHUNK + (int64_t)a.voff * LD1;
            const double* y0 = a.ypk + (int64_t)blockIdx.y * a.nkq_alloc * (BN * LDK);
            for (int it = 0; it < nit; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&full[s], (uint32_t)((2 * VH + BN * LDK) * 8));
                double* dV = sV + (size_t)s * 2 * VH;
                bulk_g2s(dV, v0 + (int64_t)it * VH, VH * 8, &full[s]);
                bulk_g2s(dV + VH, v0 + VPK_CHUNK + (int64_t)it * VH, VH * 8, &full[s]);
                bulk_g2s(sY + (size_t)s * BN * LDK, y0 + (int64_t)it * (BN * LDK), BN * LDK * 8, &full[s]);
            }
        }
        return;
    }

    // ===== DMMA consumer warps =====
    const int wm = warp / WN, wn = warp % WN;
    const int64_t rbase = m0 + wm * WTM + (lane >> 2);
    const int cbase = n0 + wn * WTN + (lane & 3) * 2;
    double acc[MI][NJ][2];
    const int fragA = (lane & 3) * LD1 + (lane >> 2);
    const int fragB = (lane >> 2) * LDK + (lane & 3);
    // columns [j0, j0 + NH) of the 8-row block i of this warp's C tile
    constexpr int NH = NJ / 2;
    auto load_half = [&](int i, int j0, double (&dst)[NH][2]) {
        const int64_t row = rbase + i * 8;
        const bool rok = row >= a.row_lo && row < a.rows;
#pragma unroll
        for (int j = 0; j < NH; ++j) {
            const int col = cbase + (j0 + j) * 8;
            const double* p = a.C + (int64_t)col * a.ldc + row;
            dst[j][0] = (rok && col < a.ncols) ? *p : 0.0;
            dst[j][1] = (rok && col + 1 < a.ncols) ? *(p + a.ldc) : 0.0;
        }
    };
    const double* v0 = sV + (wm * WTM / 64) * VH + (wm * WTM % 64) + fragA;
    const double* y0 = sY + wn * WTN * LDK + fragB;
    auto mma_steps = [&](int s, int k_lo, int k_hi) {
        const double* v = v0 + (size_t)s * 2 * VH;
        const double* y = y0 + (size_t)s * BN * LDK;
#pragma unroll
        for (int kk = k_lo; kk < k_hi; ++kk) {
            double af[MI], bf[NJ];
#pragma unroll
            for (int i = 0; i < MI; ++i) af[i] = v[kk * 4 * LD1 + i * 8];
#pragma unroll
            for (int j = 0; j < NJ; ++j) bf[j] = y[j * 8 * LDK + kk * 4];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    };
    if (DEFER) {
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 1
        for (int it = 0; it < MI; ++it) {                        // nit == MI: one 8-row block of C per k-stage, in two halves
            const int s = it % STAGES;
            double cpre[NH][2];
            load_half(it, 0, cpre);
            mbar_wait(&full[s], (it / STAGES) & 1);
            release_prev_stage(empty, it, STAGES, lane);
            mma_steps(s, 0, KC / 8);
#pragma unroll
            for (int i = 0; i < MI; ++i)                         // predicated adds: the loop over the stages stays rolled
                if (i == it) {
#pragma unroll
                    for (int j = 0; j < NH; ++j) {
                        acc[i][j][0] += cpre[j][0];
                        acc[i][j][1] += cpre[j][1];
                    }
                }
            load_half(it, NH, cpre);
            mma_steps(s, KC / 8, KC / 4);
#pragma unroll
            for (int i = 0; i < MI; ++i)
                if (i == it) {
#pragma unroll
                    for (int j = 0; j < NH; ++j) {
                        acc[i][NH + j][0] += cpre[j][0];
                        acc[i][NH + j][1] += cpre[j][1];
                    }
                }
        }
    } else {
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            double h[NH][2];
            load_half(i, 0, h);
#pragma unroll
            for (int j = 0; j < NH; ++j) { acc[i][j][0] = h[j][0]; acc[i][j][1] = h[j][1]; }
            load_half(i, NH, h);
#pragma unroll
            for (int j = 0; j < NH; ++j) { acc[i][NH + j][0] = h[j][0]; acc[i][NH + j][1] = h[j][1]; }
        }
        for (int it = 0; it < nit; ++it) {
            const int s = it % STAGES;
            mbar_wait(&full[s], (it / STAGES) & 1);
            release_prev_stage(empty, it, STAGES, lane);
            mma_steps(s, 0, KC / 4);
        }
    }
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int64_t row = rbase + i * 8;
        const bool rok = row >= a.row_lo && row < a.rows;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int col = cbase + j * 8;
            double* p = a.C + (int64_t)col * a.ldc + row;
            if (rok && col < a.ncols) *p = acc[i][j][0];
            if (rok && col + 1 < a.ncols) *(p + a.ldc) = acc[i][j][1];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// gemm_cvy_p: the 128-wide update C += V Y with CTAs that walk through `tiles_per_cta` consecutive tiles.  Same tile, same warp
// layout and the same deferred C reads as k_gemm_cvy<4, 2, true>; what changes is that the TMA producer warp runs ahead across
// tile boundaries, so the operand pipeline of a CTA does not drain between its tiles: a one-tile CTA pays launch + barrier
// set-up + the first two stage fills (about 4 of its 21 microseconds, profiles/r01_prof_cvy_ncu.txt) before its first DMMA.
// The walk is kept SHORT on purpose: under look-ahead the panel chain's kernels (high-priority stream) only get SMs when CTAs of
// the bulk update retire; fully persistent CTAs starve the chain and serialise the schedule (measured: 41.5 -> 45.4 ms).
//   tile t -> row tile t % tiles_m, column tile t / tiles_m: consecutive tiles of a CTA share the Y block (L2).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(9 * 32, 2) k_gemm_cvy_p(GemmCvyArgs a) {
    constexpr int BM = 128, BN = YT, WM = 4, WN = 2, NCW = WM * WN, STAGES = 2;
    constexpr int WTM = BM / WM, WTN = BN / WN;
    constexpr int MI = WTM / 8, NJ = WTN / 8, NH = NJ / 2;
    constexpr int VH = KC * LD1;   // doubles per 64-row x 32-col slice
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sV = reinterpret_cast<double*>(smem_raw);   // [STAGES][2][KC][LD1]
    double* sY = sV + STAGES * 2 * VH;                   // [STAGES][BN][LDK]
    uint64_t* full = reinterpret_cast<uint64_t*>(sY + STAGES * BN * LDK);
    uint64_t* empty = full + STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (wide_gate_closed(a.ctl, a.gate)) return;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NCW);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const int ntiles = a.tiles_m * a.tiles_n;
    const int t_lo = blockIdx.x * a.tiles_per_cta, t_hi = min(t_lo + a.tiles_per_cta, ntiles);

    if (warp == NCW) {
        // ===== TMA producer warp: (tile, k-stage) pairs back to back =====
        if (lane == 0) {
            int g = 0;
            for (int t = t_lo; t < t_hi; ++t) {
                const int bx = t % a.tiles_m, by = t / a.tiles_m;
                const double* v0 = a.vpk + (int64_t)(2 * bx) * VPK_CHUNK + (int64_t)a.voff * LD1;
                const double* y0 = a.ypk + (int64_t)by * a.nkq_alloc * (BN * LDK);
                for (int it = 0; it < MI; ++it, ++g) {
                    const int s = g % STAGES;
                    mbar_wait(&empty[s], ((g / STAGES) & 1) ^ 1);
                    mbar_arrive_expect_tx(&full[s], (uint32_t)((2 * VH + BN * LDK) * 8));
                    double* dV = sV + (size_t)s * 2 * VH;
                    bulk_g2s(dV, v0 + (int64_t)it * VH, VH * 8, &full[s]);
                    bulk_g2s(dV + VH, v0 + VPK_CHUNK + (int64_t)it * VH, VH * 8, &full[s]);
                    bulk_g2s(sY + (size_t)s * BN * LDK, y0 + (int64_t)it * (BN * LDK), BN * LDK * 8, &full[s]);
                }
            }
        }
        return;
    }

    // ===== DMMA consumer warps =====
    const int wm = warp / WN, wn = warp % WN;
    const int fragA = (lane & 3) * LD1 + (lane >> 2);
    const int fragB = (lane >> 2) * LDK + (lane & 3);
    const double* v0s = sV + (wm * WTM / 64) * VH + (wm * WTM % 64) + fragA;
    const double* y0s = sY + wn * WTN * LDK + fragB;
    int g = 0;
    for (int t = t_lo; t < t_hi; ++t) {
        const int bx = t % a.tiles_m, by = t / a.tiles_m;
        const int64_t rbase = (int64_t)bx * BM + wm * WTM + (lane >> 2);
        const int cbase = by * BN + wn * WTN + (lane & 3) * 2;
        double acc[MI][NJ][2];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        auto load_half = [&](int i, int j0, double (&dst)[NH][2]) {
            const int64_t row = rbase + i * 8;
            const bool rok = row >= a.row_lo && row < a.rows;
#pragma unroll
            for (int j = 0; j < NH; ++j) {
                const int col = cbase + (j0 + j) * 8;
                const double* p = a.C + (int64_t)col * a.ldc + row;
                dst[j][0] = (rok && col < a.ncols) ? *p : 0.0;
                dst[j][1] = (rok && col + 1 < a.ncols) ? *(p + a.ldc) : 0.0;
            }
        };
        auto mma_steps = [&](int s, int k_lo, int k_hi) {
            const double* v = v0s + (size_t)s * 2 * VH;
            const double* y = y0s + (size_t)s * BN * LDK;
#pragma unroll
            for (int kk = k_lo; kk < k_hi; ++kk) {
                double af[MI], bf[NJ];
#pragma unroll
                for (int i = 0; i < MI; ++i) af[i] = v[kk * 4 * LD1 + i * 8];
#pragma unroll
                for (int j = 0; j < NJ; ++j) bf[j] = y[j * 8 * LDK + kk * 4];
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        };
#pragma unroll 1
        for (int it = 0; it < MI; ++it, ++g) {                   // one 8-row block of C per k-stage, in two halves
            const int s = g % STAGES;
            double cpre[NH][2];
            load_half(it, 0, cpre);
            mbar_wait(&full[s], (g / STAGES) & 1);
            release_prev_stage(empty, g, STAGES, lane);
            mma_steps(s, 0, KC / 8);
#pragma unroll
            for (int i = 0; i < MI; ++i)
                if (i == it) {
#pragma unroll
                    for (int j = 0; j < NH; ++j) {
                        acc[i][j][0] += cpre[j][0];
                        acc[i][j][1] += cpre[j][1];
                    }
                }
            load_half(it, NH, cpre);
            mma_steps(s, KC / 8, KC / 4);
#pragma unroll
            for (int i = 0; i < MI; ++i)
                if (i == it) {
#pragma unroll
                    for (int j = 0; j < NH; ++j) {
                        acc[i][NH + j][0] += cpre[j][0];
                        acc[i][NH + j][1] += cpre[j][1];
                    }
                }
        }
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const int64_t row = rbase + i * 8;
            const bool rok = row >= a.row_lo && row < a.rows;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int col = cbase + j * 8;
                double* p = a.C + (int64_t)col * a.ldc + row;
                if (rok && col < a.ncols) *p = acc[i][j][0];
                if (rok && col + 1 < a.ncols) *(p + a.ldc) = acc[i][j][1];
            }
        }
    }
    // the last stage of the last tile is never released: nobody waits for it
}

constexpr int WP = 128;            // wide panel width
constexpr int WLD = WP + 1;        // leading dimension of the row-major 128 x 128 work matrices in shared memory

// ------------------------------------------------------------------------------------------------
// Inverse of an upper triangular 32 x 32 block by ONE warp (lane = column of the inverse, back substitution in registers;
// every lane runs the same 496 multiply-adds, the entries of R are broadcast loads).  R: row-major, leading dimension ldr;
// dinv = 1 / diag(R) or null; D: row-major 32 x 32 with leading dimension LDD (zeros below the diagonal).
// ------------------------------------------------------------------------------------------------
constexpr int LDD = 33;   // leading dimension of an inverted 32 x 32 block in shared memory
__device__ __forceinline__ void triu_inv32_warp(const double* R, int ldr, const double* dinv, double* D, int lane) {
    double x[32];
#pragma unroll
    for (int i = 31; i >= 0; --i) {
        double s = i == lane ? 1.0 : 0.0;
#pragma unroll
        for (int p = i + 1; p < 32; ++p) s -= R[i * ldr + p] * x[p];       // x[p] == 0 for p > lane
        const double di = dinv ? dinv[i] : 1.0 / R[i * ldr + i];
        x[i] = i <= lane ? s * di : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) D[i * LDD + lane] = x[i];
}

// ------------------------------------------------------------------------------------------------
// In-place inverse of the upper triangle of A (row-major, leading dimension WLD, 128 x 128), 16 warps: the four 32 x 32
// diagonal blocks by triu_inv32_warp, then two levels of X12 = -X11 (R12 X22) on the fp64 tensor pipe (the scalar form is
// bound by its two shared-memory loads per multiply-add).  dinv = 1 / diag or null (diagonal read from A).  The strict lower
// triangle of A must be ZERO inside the diagonal 32 x 32 blocks (the tensor-pipe tiles on the diagonal read it).
// T: scratch of 4 * 32 * LDD doubles (diagonal blocks), T2: scratch of 64 * 65 doubles (R12 X22).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void triu_inv128_mma(double* A, const double* dinv, double* T, double* T2, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    if (warp < 4) triu_inv32_warp(A + (32 * warp) * WLD + 32 * warp, WLD, dinv ? dinv + 32 * warp : nullptr, T + warp * 32 * LDD, lane);
    __syncthreads();
    for (int e = tid; e < 4096; e += 512) {
        const int b = e >> 10, i = (e >> 5) & 31, c = e & 31;
        if (i <= c) A[(32 * b + i) * WLD + 32 * b + c] = T[b * 32 * LDD + i * LDD + c];
    }
    __syncthreads();
#pragma unroll 1
    for (int bs = 32; bs < WP; bs *= 2) {
        const int nt = bs / 8, ntile = nt * nt, npairs = WP / (2 * bs), ld2 = bs + 1;
        // Tm = R12 X22 (X22 upper triangular: k <= column)
        for (int idx = warp; idx < npairs * ntile; idx += 16) {
            const int pi = idx / ntile, tt = idx % ntile, ti = tt / nt, tj = tt % nt, o = pi * 2 * bs;
            const double* pa = A + (o + 8 * ti + (lane >> 2)) * WLD + o + bs + (lane & 3);          // R12(i, k)
            const double* pb = A + (o + bs + (lane & 3)) * WLD + o + bs + 8 * tj + (lane >> 2);      // X22(k, j)
            double c0 = 0.0, c1 = 0.0;
            for (int k4 = 0; k4 < 2 * (tj + 1); ++k4) dmma(c0, c1, pa[4 * k4], pb[4 * k4 * WLD]);
            double* pt = T2 + pi * bs * ld2 + (8 * ti + (lane >> 2)) * ld2 + 8 * tj + 2 * (lane & 3);
            pt[0] = c0;
            pt[1] = c1;
        }
        __syncthreads();
        // X12 = -X11 Tm (X11 upper triangular: k >= row)
        for (int idx = warp; idx < npairs * ntile; idx += 16) {
            const int pi = idx / ntile, tt = idx % ntile, ti = tt / nt, tj = tt % nt, o = pi * 2 * bs;
            const double* pa = A + (o + 8 * ti + (lane >> 2)) * WLD + o + (lane & 3);               // X11(i, k)
            const double* pb = T2 + pi * bs * ld2 + (lane & 3) * ld2 + 8 * tj + (lane >> 2);         // Tm(k, j)
            double c0 = 0.0, c1 = 0.0;
            for (int k4 = 2 * ti; k4 < 2 * nt; ++k4) dmma(c0, c1, pa[4 * k4], pb[4 * k4 * ld2]);
            double* pc = A + (o + 8 * ti + (lane >> 2)) * WLD + o + bs + 8 * tj + 2 * (lane & 3);
            pc[0] = -c0;
            pc[1] = -c1;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// tinv:  Linv = (I + stril(S))^{-1},  S = first NBP ext columns of the reduced Wext.
//   With |v|^2 = 2 the compact-WY factor obeys T^{-1} = I + striu(V'V), so Linv == T'.
//   One CTA; 8x8 diagonal blocks by forward substitution, then log2(NBP/8) merge levels
//   X21 = -X22 (L21 X11).
// ------------------------------------------------------------------------------------------------
// in-place inversion of the unit lower-triangular L (smem, element (i,j) at j*LDL + i, strict lower part filled,
// rest zero) -> L holds (I + stril)^{-1} including the unit diagonal; T = scratch of >= max(NBP*9, NBP*NBP/4... 4224) doubles
template <int NBP>
__device__ __forceinline__ void tinv_core(double* L, double* T, int tid, int nthreads) {
    constexpr int LDL = NBP + 1;
    // diagonal 8x8 blocks: X = (I + N)^{-1} by forward substitution; thread = one column of one block
    constexpr int NDB = NBP / 8;
    if (tid < NDB * 8) {
        const int d = tid >> 3, j = tid & 7;
        const double* Ld = L + (d * 8) * LDL + d * 8;
        double* X = T + d * 72;   // (i,j) at j*9 + i
#pragma unroll
        for (int i = 0; i < 8; ++i) X[j * 9 + i] = (i == j) ? 1.0 : 0.0;
#pragma unroll
        for (int i = 1; i < 8; ++i) {
            double accv = 0.0;
            for (int k = 0; k < i; ++k) accv += Ld[k * LDL + i] * X[j * 9 + k];
            if (i > j) X[j * 9 + i] = -accv;
        }
    }
    __syncthreads();
    for (int e = tid; e < NDB * 64; e += nthreads) {
        const int d = e / 64, r = e % 64, i = r % 8, j = r / 8;
        L[(d * 8 + j) * LDL + d * 8 + i] = T[d * 72 + j * 9 + i];
    }
    __syncthreads();
    // merge levels: X21 = -X22 (L21 X11) for every pair of adjacent inverted blocks
    for (int bs = 8; bs < NBP; bs *= 2) {
        const int npairs = NBP / (2 * bs);
        for (int e = tid; e < npairs * bs * bs; e += nthreads) {
            const int p = e / (bs * bs), r = e % (bs * bs), i = r % bs, j = r / bs, o = p * 2 * bs;
            double sacc = 0.0;
            for (int k = j; k < bs; ++k) sacc += L[(o + k) * LDL + o + bs + i] * L[(o + j) * LDL + o + k];
            T[p * bs * bs + j * bs + i] = sacc;
        }
        __syncthreads();
        for (int e = tid; e < npairs * bs * bs; e += nthreads) {
            const int p = e / (bs * bs), r = e % (bs * bs), i = r % bs, j = r / bs, o = p * 2 * bs;
            double sacc = 0.0;
            for (int k = 0; k <= i; ++k) sacc += L[(o + bs + k) * LDL + o + bs + i] * T[p * bs * bs + j * bs + k];
            L[(o + j) * LDL + o + bs + i] = -sacc;
        }
        __syncthreads();
    }
}

// grid.x > 1: a batch, CTA g inverts the block at Ws + g * ws_stride into Linv + g * NBP * NBP
template <int NBP>
__global__ void __launch_bounds__(512, 1) k_tinv(const double* __restrict__ Ws, double* __restrict__ Linv, int64_t ws_stride = 0) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    Ws += (int64_t)blockIdx.x * ws_stride;
    Linv += (int64_t)blockIdx.x * NBP * NBP;
    if (NBP == WP) {
        // (I + stril(S))^{-1} = ((I + striu(S'))^{-1})': row r of U = I + striu(S') is column r of S below the diagonal, i.e. a
        // contiguous run of Ws, and row j of U^{-1} is column j of the result: both transfers are linear in memory
        double* A = reinterpret_cast<double*>(smem_raw);   // [128][WLD] row-major
        double* T = A + WP * WLD;                            // 4 * 32 * LDD
        double* T2 = T + 4 * 32 * LDD;                       // 64 * 65
        for (int e = tid; e < WP * WP; e += 512) {
            const int r = e >> 7, cc = e & (WP - 1);
            A[r * WLD + cc] = cc > r ? Ws[e] : (cc == r ? 1.0 : 0.0);   // zeros below: the 8 x 8 tiles on the diagonal read them
        }
        __syncthreads();
        triu_inv128_mma(A, nullptr, T, T2, tid);
        for (int e = tid; e < WP * WP; e += 512) {
            const int j = e >> 7, i = e & (WP - 1);
            Linv[e] = i >= j ? A[j * WLD + i] : 0.0;
        }
    } else {
        constexpr int LDL = NBP + 1;
        double* L = reinterpret_cast<double*>(smem_raw);   // [NBP][LDL], element (i,j) at j*LDL + i
        double* T = L + NBP * LDL;                          // scratch, 4 * 32 * 33 doubles
        for (int e = tid; e < NBP * NBP; e += blockDim.x) {
            const int i = e % NBP, j = e / NBP;
            L[j * LDL + i] = (i > j) ? Ws[e] : 0.0;
        }
        __syncthreads();
        tinv_core<NBP>(L, T, tid, blockDim.x);
        for (int e = tid; e < NBP * NBP; e += blockDim.x) {
            const int i = e % NBP, j = e / NBP;
            Linv[e] = (i >= j) ? L[j * LDL + i] : 0.0;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// mid32: the whole middle of a 32-wide block update in one launch (the inner-panel updates sit on the
// critical path of the panel chain, where every launch costs):  split-K reduction of the Gram block and of
// this CTA's 32 W columns (fixed order), T' = (I + stril(S))^{-1}, Y = -T'W in the packed ypk layout.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1) k_mid32(const double* __restrict__ Wp, int64_t pstride, int nsplit, int na,
                                                  double* __restrict__ ypk, double* __restrict__ linv_out, int trans) {
    constexpr int NBP = 32, LDL = 33;
    __shared__ double L[NBP * LDL];
    __shared__ double T[1024];
    __shared__ double sW[YCOLS * NBP];
    const int tid = threadIdx.x;
    const int c0 = blockIdx.x * YCOLS;
    // elements 0..1023: Gram block (i, j); 1024..2047: W tile (k, j)
    for (int e = tid; e < 2048; e += 512) {
        const bool isS = e < 1024;
        const int r = e & 1023, i = r % NBP, j = r / NBP;
        const bool need = isS ? (i > j) : (c0 + j < na);
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        if (need) {
            const double* src = Wp + (isS ? (int64_t)r : (int64_t)(NBP + c0) * NBP + r);
            int p = 0;
            for (; p + 4 <= nsplit; p += 4) {
                s0 += src[(int64_t)p * pstride];
                s1 += src[(int64_t)(p + 1) * pstride];
                s2 += src[(int64_t)(p + 2) * pstride];
                s3 += src[(int64_t)(p + 3) * pstride];
            }
            for (; p < nsplit; ++p) s0 += src[(int64_t)p * pstride];
        }
        const double v = (s0 + s1) + (s2 + s3);
        if (isS) L[j * LDL + i] = v; else sW[r] = v;
    }
    __syncthreads();
    tinv_core<NBP>(L, T, tid, 512);
    if (blockIdx.x == 0)   // keep T' for a later block update with the same V (look-ahead part (b))
        for (int e = tid; e < NBP * NBP; e += 512) {
            const int i = e % NBP, j = e / NBP;
            linv_out[e] = (i >= j) ? L[j * LDL + i] : 0.0;
        }
    // Y(i, col) = -sum_{k<=i} Linv(i,k) W(k,col);  thread = (i, two columns)
    const int i = tid % NBP, jg = tid / NBP;   // jg in 0..15
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int j = jg * 2 + h;
        double acc = 0.0;
        if (!trans) for (int k = 0; k <= i; ++k) acc += L[k * LDL + i] * sW[j * NBP + k];        // Y = -T' W  (Q' C)
        else for (int k = i; k < NBP; ++k) acc += L[i * LDL + k] * sW[j * NBP + k];               // Y = -T W   (Q C)
        const int col = c0 + j;
        ypk[(int64_t)(col / YT) * (YT * LDK) + (col % YT) * LDK + i] = (col < na) ? -acc : 0.0;   // NKQ == 1
    }
}

// ------------------------------------------------------------------------------------------------
// ymake:  Y(NBP x na) = -Linv * W   (W = ext columns [NBP, NBP+na) of the reduced Wext), written in
//   the packed layout gemm_cvy stages with one bulk copy:  ypk[col/64][k/32][col%64][LDK].
//   CTA = YCOLS columns; thread = (row i, a group of the columns).
// ------------------------------------------------------------------------------------------------
template <int NBP>
__global__ void __launch_bounds__(256, 1) k_ymake(const double* __restrict__ Ws, int woff, int na, const double* __restrict__ Linv,
                                                  double* __restrict__ ypk, int trans) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sL = reinterpret_cast<double*>(smem_raw);   // [NBP][NBP] col-major
    double* sW = sL + NBP * NBP;                         // [YCOLS][NBP]
    const int tid = threadIdx.x;
    const int c0 = blockIdx.x * YCOLS;
    const int nc = min(YCOLS, na - c0);
    for (int e = tid; e < NBP * NBP; e += blockDim.x) sL[e] = Linv[e];
    for (int e = tid; e < YCOLS * NBP; e += blockDim.x) {
        const int j = e / NBP;
        sW[e] = (j < nc) ? Ws[(int64_t)(woff + c0) * NBP + e] : 0.0;   // woff = ext column where W starts (NBP, or 0 when T is reused)
    }
    __syncthreads();
    constexpr int TPR = 256 / NBP;           // threads per row (2 for 128, 8 for 32)
    constexpr int CPT = YCOLS / TPR;         // columns per thread
    constexpr int NKQ = NBP / KC;
    const int i = tid % NBP, jh = tid / NBP;
    double acc[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[j] = 0.0;
    if (!trans) {
        for (int k = 0; k <= i; ++k) {            // Linv = T' is lower triangular: Y = -T' W (the block form of Q' C)
            const double l = sL[k * NBP + i];
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[j] += l * sW[(jh * CPT + j) * NBP + k];
        }
    } else {
        for (int k = i; k < NBP; ++k) {           // Y = -T W = -Linv' W (the block form of Q C)
            const double l = sL[i * NBP + k];
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[j] += l * sW[(jh * CPT + j) * NBP + k];
        }
    }
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int col = c0 + jh * CPT + j;     // columns beyond na get zeros (the tile is copied whole)
        ypk[((int64_t)(col / YT) * NKQ + i / KC) * (YT * LDK) + (col % YT) * LDK + (i % KC)] = (col < na) ? -acc[j] : 0.0;
    }
}

// ------------------------------------------------------------------------------------------------
// panel:  cooperative, persistent factorisation of an mp x ncols (ncols <= IB) panel.
//   Each CTA keeps a slab of rows in shared memory for the whole kernel.  One grid-wide
//   reduction per column: the same pass that applies reflector j also accumulates
//   x'a_c (x = next pivot column, rows > j) for every remaining column c, so that after the exchange
//   every CTA can form   s=|x|, alpha, f   and   w_c = v'a_c = f (x'a_c - alpha a_c[j])   locally
//   (S:129-131 and S:208 in one reduction).
//   The exchange goes through L2 in self-validating cells (the NCCL "LL" idea): every 8-byte word
//   carries 32 data bits and a 32-bit tag unique to (launch, column) -> no fence, no atomic, no
//   barrier.  Two levels keep the traffic and the number of pollers small: each CTA publishes its
//   partials; the owner warp of column c (CTA c % G) sums the G partials in CTA order (deterministic)
//   and publishes one total; every CTA then polls at most IB totals and IB pivot-row cells.
//   Also writes the packed V block (vpk) for the GEMMs.
// ------------------------------------------------------------------------------------------------
struct PanelArgs {
    double* P;            // panel top-left (row = pivot row of column 0)
    int64_t ldp;
    int64_t mp;           // rows
    int ncols;            // active columns (<= IB)
    double* alpha;        // alpha[0:ncols]
    double* vpk;          // packed V of the outer panel (may be null)
    int voff;             // first packed column of this sub-panel
    int64_t vtop;         // window rows above the panel top (zero-filled in vpk)
    int64_t vrows;        // total window rows incl. padding (zero-filled below vtop+mp)
    int rows_per_cta;
    int lds;              // slab leading dimension (>= rows_per_cta rounded up to 4, == 4 mod 8: conflict-free DMMA fragments)
    unsigned long long* cells;   // [IB steps][(G + 2) * IB cells][2 words]
    uint32_t epoch;       // tags epoch+1 .. epoch+IB belong to this launch
    int backoff;          // ns to sleep between polls of a cell that is not there yet (0 = spin)
    unsigned long long* cells2;  // exchange cells of the CholeskyQR2 fast path: 2 x [(G+1) x 528] + 1088 cells
    int fast;             // 1: try CholeskyQR2 + Householder reconstruction first (3 exchanges per panel instead of 32)
    int* fast_stats;      // optional [2]: number of panels done by the fast path / by the column-wise fallback
    int levels;           // 2: owner warp gathers the partials and publishes a total; 1: every CTA gathers all partials itself
    long long* trace;     // optional clock64() stamps [gridDim.x][IB][8] (debugging / tuning); null = off
    const WideCtl* ctl;   // speculative panel chain: skip when a panel below `gate` was refused (may be null)
    int gate;
};

__device__ __forceinline__ void ll_store(unsigned long long* cell, double v, uint32_t tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    const unsigned long long t = (unsigned long long)tag << 32;
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(cell), "l"((b & 0xffffffffull) | t), "l"((b >> 32) | t) : "memory");
}
__device__ __forceinline__ void ll_peek(const unsigned long long* cell, unsigned long long& w0, unsigned long long& w1) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(cell) : "memory");
}
__device__ __forceinline__ double ll_finish(const unsigned long long* cell, unsigned long long w0, unsigned long long w1, uint32_t tag,
                                            int backoff) {
    while ((uint32_t)(w0 >> 32) != tag || (uint32_t)(w1 >> 32) != tag) {
        if (backoff > 0) __nanosleep(backoff);
        ll_peek(cell, w0, w1);
    }
    return __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
}
__device__ __forceinline__ double ll_wait(const unsigned long long* cell, uint32_t tag, int backoff) {
    unsigned long long w0, w1;
    ll_peek(cell, w0, w1);
    return ll_finish(cell, w0, w1, tag, backoff);
}

constexpr int PANEL_MAXG = 160;   // max CTAs of the panel kernel (owner gather: 5 cells per lane)
constexpr int PNW = PANEL_THREADS / 32;

__global__ void __launch_bounds__(PANEL_THREADS, 1) k_panel(PanelArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* S = reinterpret_cast<double*>(smem_raw);   // [IB][lds]
    __shared__ double tot[2][IB];
    __shared__ double pv[2][IB];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, cta = blockIdx.x;
    if (wide_gate_closed(a.ctl, a.gate)) return;   // grid-uniform: every CTA reads the same word before any exchange
    const int64_t row0 = (int64_t)cta * a.rows_per_cta;
    const int nr = (int)max((int64_t)0, min((int64_t)a.rows_per_cta, a.mp - row0));
    const int lds = a.lds, nc = a.ncols;
    const size_t step_words = ((size_t)G * IB + 2 * IB) * 2;
    auto pcell = [&](int step, int g, int c) { return a.cells + (size_t)step * step_words + ((size_t)g * IB + c) * 2; };
    auto tcell = [&](int step, int c) { return a.cells + (size_t)step * step_words + ((size_t)G * IB + c) * 2; };
    auto vcell = [&](int step, int c) { return a.cells + (size_t)step * step_words + ((size_t)G * IB + IB + c) * 2; };
    auto clampi = [&](int64_t x) { return x < 0 ? 0 : (x > nr ? nr : (int)x); };

    // load slab (coalesced along rows); rows [nr, nr4) are zero so that the DMMAs can run on 8-row tiles
    const int nr4 = (nr + 7) & ~7;
    for (int c = warp; c < nc; c += PNW)
        for (int r = lane; r < nr4; r += 32) S[c * lds + r] = r < nr ? a.P[(int64_t)c * a.ldp + row0 + r] : 0.0;
    __syncthreads();

    // ============================================================================================
    // Fast path: CholeskyQR2 + Householder reconstruction (Ballard, Demmel, Grigori, Jacquelin, Nguyen,
    // Solomonik 2014).  P = Q Rt by two Cholesky-QR passes (2 exchanges: the 32x32 Gram matrices), then the
    // unique Householder representation of that QR is recovered from the LU factorisation of E - Q S
    // (S_jj = -sign(Q_jj^(j)) chosen on the fly, pivots U_jj = 1 + |Q_jj^(j)| = tau_j): the top 32x32 block on
    // CTA 0, one more exchange for its frozen rows, and a row-local triangular solve everywhere else.  In the
    // reference's storage: v_ij = W_ij^(j) / sqrt(U_jj) (i > j), v_jj = -S_j sqrt(U_jj), alpha_j = S_j Rt_jj,
    // R_ij = S_i Rt_ij.  Same reflectors as S:127-135 up to rounding (verified against the oracle), but it
    // squares the panel's condition number on the way: if a Cholesky pivot is not positive, the first factor's
    // diagonal spans more than 250 (1e5 with substitution solves), or Q1'Q1 is further than 1/4 from I (i.e. the second pass could not restore
    // orthogonality to O(eps)), the slab is reloaded and the column-by-column path below runs instead
    // (that also reproduces the reference's NaN behaviour for zero columns).  The decision is taken from
    // identical data on every CTA, so it is grid-uniform without another exchange.
    // ============================================================================================
    bool done = false;
    if (a.fast && nc == IB && a.mp >= 2 * IB && a.rows_per_cta >= IB) {
        constexpr int NGP = IB * (IB + 1) / 2;   // 528 pairs (i >= j)
        constexpr int LDG = IB + 1;
        __shared__ double Gm[IB * LDG], Wt[IB * LDG];
        __shared__ __align__(16) double R1[IB * IB], R2[IB * IB];   // Cholesky factors, rows 16-byte aligned (paired loads)
        __shared__ double rinv[IB], Sg[IB], Ud[IB], rsq[IB], cl[IB];
        __shared__ double Dv[4 * 64], Wn[6 * 64];   // DMMA triangular solve: inverses of the 8x8 diagonal blocks, -R_ab * inv(R_bb)
        __shared__ int bad;
        const size_t reg = (size_t)(G + 1) * NGP;
        auto c2p = [&](int e, int g, int t) { return a.cells2 + ((size_t)e * reg + (size_t)g * NGP + t) * 2; };
        auto c2t = [&](int e, int t) { return a.cells2 + ((size_t)e * reg + (size_t)G * NGP + t) * 2; };
        auto c2u = [&](int idx) { return a.cells2 + (2 * reg + idx) * 2; };   // 1024 frozen-row cells, then Ud[32], Sg[32]
        if (tid == 0) bad = 0;

        // Gram matrix of the slab -> all-CTA sum in Gm (both triangles); partials summed in CTA order
        auto gram_exchange = [&](int e, uint32_t tag) {
            // partial Gram of the slab on the fp64 tensor pipe: the lower 16x16 tiles (0,0), (1,0), (1,1) of S'S, each
            // by NKG warps that split the slab's 4-row steps; the NKG partials of a tile are added into Gm in group
            // order (fixed -> deterministic).  Both fragments of a DMMA are slab columns, k = slab rows ("TN" like gemm_vta).
            {
                constexpr int NKG = 5;
                const int tile = warp % 3, kg = warp / 3;
                const int ti = tile == 0 ? 0 : 1, tj = tile == 2 ? 1 : 0;
                double acc[2][2][2];
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int v = 0; v < 2; ++v) acc[u][v][0] = acc[u][v][1] = 0.0;
                if (warp < 3 * NKG) {
                    const int ks = nr4 >> 2;
                    const int k_lo = (ks * kg) / NKG, k_hi = (ks * (kg + 1)) / NKG;
                    const double* pa = S + (ti * 16 + (lane >> 2)) * lds + (lane & 3);
                    const double* pb = S + (tj * 16 + (lane >> 2)) * lds + (lane & 3);
#pragma unroll 2
                    for (int k = k_lo; k < k_hi; ++k) {
                        const double a0 = pa[4 * k], a1 = pa[8 * lds + 4 * k];
                        const double b0 = pb[4 * k], b1 = pb[8 * lds + 4 * k];
                        dmma(acc[0][0][0], acc[0][0][1], a0, b0);
                        dmma(acc[0][1][0], acc[0][1][1], a0, b1);
                        dmma(acc[1][0][0], acc[1][0][1], a1, b0);
                        dmma(acc[1][1][0], acc[1][1][1], a1, b1);
                    }
                }
                for (int g = 0; g < NKG; ++g) {
                    if (warp < 3 * NKG && kg == g) {
#pragma unroll
                        for (int u = 0; u < 2; ++u)
#pragma unroll
                            for (int v = 0; v < 2; ++v) {
                                double* o = Gm + (ti * 16 + u * 8 + (lane >> 2)) * LDG + tj * 16 + v * 8 + 2 * (lane & 3);
                                if (g == 0) { o[0] = acc[u][v][0]; o[1] = acc[u][v][1]; }
                                else { o[0] += acc[u][v][0]; o[1] += acc[u][v][1]; }
                            }
                    }
                    __syncthreads();
                }
                for (int x = tid; x < IB * IB; x += PANEL_THREADS) {
                    const int i = x / IB, j = x % IB;
                    if (j > i) continue;
                    ll_store(c2p(e, cta, i * (i + 1) / 2 + j), Gm[i * LDG + j], tag);
                }
            }
            for (int q = warp; cta + q * G < NGP; q += PNW) {   // owner of pair t = cta + q G
                const int t = cta + q * G;
                unsigned long long w0[PANEL_MAXG / 32], w1[PANEL_MAXG / 32];
#pragma unroll
                for (int u = 0; u < PANEL_MAXG / 32; ++u)
                    if (lane + 32 * u < G) ll_peek(c2p(e, lane + 32 * u, t), w0[u], w1[u]);
                double sum = 0.0;
#pragma unroll
                for (int u = 0; u < PANEL_MAXG / 32; ++u)
                    if (lane + 32 * u < G) sum += ll_finish(c2p(e, lane + 32 * u, t), w0[u], w1[u], tag, a.backoff);
                sum = warp_sum(sum);
                if (lane == 0) ll_store(c2t(e, t), sum, tag);
            }
            {   // totals: both cells of a thread are requested before either is waited for (one L2 round trip, not two)
                static_assert(IB * IB == 2 * PANEL_THREADS, "two cells per thread");
                unsigned long long w0[2], w1[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int x = tid + u * PANEL_THREADS, i = x / IB, j = x % IB;
                    if (j <= i) ll_peek(c2t(e, i * (i + 1) / 2 + j), w0[u], w1[u]);
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int x = tid + u * PANEL_THREADS, i = x / IB, j = x % IB;
                    if (j <= i) {
                        const double v = ll_finish(c2t(e, i * (i + 1) / 2 + j), w0[u], w1[u], tag, a.backoff);
                        Gm[i * LDG + j] = v;
                        Gm[j * LDG + i] = v;
                    }
                }
            }
            __syncthreads();
        };
        // x[j+1 ..] -= q * Rrow[j+1 ..] with paired (16-byte) broadcast loads; every condition folds once j is unrolled
#define DHQR_ROW_AXPY(x, q, Rrow, j)                                                          \
        _Pragma("unroll") for (int k_ = 0; k_ < IB; k_ += 2) {                                \
            if (k_ > (j)) {                                                                   \
                const double2 rr_ = *reinterpret_cast<const double2*>((Rrow) + k_);           \
                x[k_] -= (q) * rr_.x;                                                         \
                x[k_ + 1] -= (q) * rr_.y;                                                     \
            } else if (k_ == (j)) {                                                           \
                x[k_ + 1] -= (q) * (Rrow)[k_ + 1];                                            \
            }                                                                                 \
        }
        // rsqrt(double) without the library's slow-path branch (keeps the step a single basic block the scheduler can interleave
        // with the rank-1 updates): MUFU seed + one cubic step, same operations as the fast path of rsqrt()
        auto rsqrt_nb = [](double d) {
            double y0;
            asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
            const double t = y0 * y0;
            const double e = fma(d, -t, 1.0);
            const double pq = fma(e, 0.375, 0.5);
            const double q = y0 * e;
            return fma(pq, q, y0);
        };
        // upper Cholesky factor of Gm -> Rout ([IB][IB], zero below the diagonal), 1/diag -> rinv; flags a non-positive /
        // non-finite pivot.  One warp, column `lane` of the trailing matrix in registers, row j broadcast through Rout: a step
        // is a dependent chain (rsqrt -> scale -> update of the next pivot), so more threads would only add barriers.
        // Software-pipelined: the next pivot is updated first and its rsqrt is issued before the remaining rank-1 updates
        // of the step, so that the one warp's in-order issue overlaps the two.
        auto chol = [&](double* Rout) {
            if (warp == 0) {
                double g[IB];
#pragma unroll
                for (int i = 0; i < IB; ++i) g[i] = Gm[i * LDG + lane];
                double d = __shfl_sync(0xffffffffu, g[0], 0);
                double ri = rsqrt_nb(d);
#pragma unroll
                for (int j = 0; j < IB; ++j) {
                    if (lane == 0) {
                        if (!(d > 0.0) || !(d < 1e300)) bad = 1;
                        rinv[j] = ri;
                    }
                    const double r = lane == j ? d * ri : g[j] * ri;   // R(j, lane)
                    Rout[j * IB + lane] = lane >= j ? r : 0.0;
                    double dn = 1.0, rin = 1.0;
                    if (j + 1 < IB) {
                        const double rj1 = __shfl_sync(0xffffffffu, r, j + 1);
                        if (j + 1 <= lane) g[j + 1] -= rj1 * r;
                        dn = __shfl_sync(0xffffffffu, g[j + 1], j + 1);
                        rin = rsqrt_nb(dn);
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = j + 2; i < IB; ++i) {
                        const double rji = Rout[j * IB + i];            // broadcast
                        if (i <= lane) g[i] -= rji * r;
                    }
                    d = dn;
                    ri = rin;
                }
            }
            __syncthreads();
        };
        // Triangular solve on the fp64 tensor pipe: slab rows [8 tile_lo, nr) <- rows * Rm^{-1}, Rm upper triangular ([IB][IB]),
        // dgi = 1 / diag(Rm).  Blocked by 8 columns: X'_b = X_b inv(R_bb) - sum_{a<b} X'_a (R_ab inv(R_bb)); the 8x8 diagonal
        // inverses and the 6 products are formed once per call (Dv, Wn), the B fragments live in registers, and one warp
        // owns an 8-row tile (A fragments = slab columns, the finished block goes back through the slab to change layout).
        // The substitution above is bound by the shared-memory -> register bandwidth of its 496 broadcast operands per row.
        auto trsm_dmma = [&](const double* Rm, const double* dgi, int tile_lo) {
            if (tid < 32) {
                const int b = tid >> 3, c = tid & 7;
                const double* Rb = Rm + (8 * b) * IB + 8 * b;
                double inv[8];
#pragma unroll
                for (int i = 7; i >= 0; --i) {
                    double sacc = i == c ? 1.0 : 0.0;
#pragma unroll
                    for (int j = i + 1; j < 8; ++j) sacc -= Rb[i * IB + j] * inv[j];
                    inv[i] = i <= c ? sacc * dgi[8 * b + i] : 0.0;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) Dv[b * 64 + i * 8 + c] = inv[i];
            }
            __syncthreads();
            if (tid < 6 * 64) {
                const int pr = tid >> 6, i = (tid >> 3) & 7, j = tid & 7;
                const int ba = pr < 3 ? 0 : (pr < 5 ? 1 : 2), bb = pr < 3 ? pr + 1 : (pr < 5 ? pr - 1 : 3);
                double sacc = 0.0;
#pragma unroll
                for (int k = 0; k < 8; ++k) sacc += Rm[(8 * ba + i) * IB + 8 * bb + k] * Dv[bb * 64 + k * 8 + j];
                Wn[pr * 64 + i * 8 + j] = -sacc;
            }
            __syncthreads();
            double bd[4][2], bw[6][2];
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int h = 0; h < 2; ++h) bd[b][h] = Dv[b * 64 + (4 * h + (lane & 3)) * 8 + (lane >> 2)];
#pragma unroll
            for (int pr = 0; pr < 6; ++pr)
#pragma unroll
                for (int h = 0; h < 2; ++h) bw[pr][h] = Wn[pr * 64 + (4 * h + (lane & 3)) * 8 + (lane >> 2)];
            const int ntiles = nr4 >> 3;
            for (int tile = tile_lo + warp; tile < ntiles; tile += PNW) {
                const double* pa = S + (lane & 3) * lds + tile * 8 + (lane >> 2);
                double* pc = S + (2 * (lane & 3)) * lds + tile * 8 + (lane >> 2);
                double xa[3][2];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const double o0 = pa[(8 * b) * lds], o1 = pa[(8 * b + 4) * lds];
                    double c0 = 0.0, c1 = 0.0;
                    dmma(c0, c1, o0, bd[b][0]);
                    dmma(c0, c1, o1, bd[b][1]);
#pragma unroll
                    for (int a2 = 0; a2 < b; ++a2) {
                        const int pr = a2 == 0 ? b - 1 : (a2 == 1 ? b + 1 : 5);   // (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
                        dmma(c0, c1, xa[a2][0], bw[pr][0]);
                        dmma(c0, c1, xa[a2][1], bw[pr][1]);
                    }
                    __syncwarp();   // every lane has read X_b in the A layout before it is overwritten in the C layout
                    pc[(8 * b) * lds] = c0;
                    pc[(8 * b + 1) * lds] = c1;
                    if (b < 3) {
                        __syncwarp();
                        xa[b][0] = pa[(8 * b) * lds];
                        xa[b][1] = pa[(8 * b + 4) * lds];
                    }
                }
            }
            __syncthreads();
        };
        // slab <- slab * R^{-1}  (row-local forward substitution; one row per thread)
        auto trsm = [&](const double* R) {
            for (int r = tid; r < nr; r += PANEL_THREADS) {
                double x[IB];
#pragma unroll
                for (int k = 0; k < IB; ++k) x[k] = S[k * lds + r];
#pragma unroll
                for (int j = 0; j < IB; ++j) {
                    const double q = x[j] * rinv[j];
                    x[j] = q;
                    DHQR_ROW_AXPY(x, q, R + j * IB, j)
                    asm volatile("" ::: "memory");   // keeps ptxas from hoisting all the LDS (it spills 4 KB/thread otherwise)
                }
#pragma unroll
                for (int k = 0; k < IB; ++k) S[k * lds + r] = x[k];
            }
            __syncthreads();
        };

        const uint32_t ftag = a.epoch + IB + 1;
        long long* ftr = a.trace ? a.trace + (size_t)cta * IB * 8 : nullptr;
        const long long ft0 = clock64();
        auto stamp = [&](int k) { if (ftr && tid == 0) ftr[k] = clock64() - ft0; };
        gram_exchange(0, ftag);
        stamp(1);
        chol(R1);
        stamp(2);
        if (tid == 0 && !bad) {   // conditioning guard on the first factor
            double dmin = R1[0], dmax = R1[0];
            for (int j = 1; j < IB; ++j) { dmin = fmin(dmin, R1[j * IB + j]); dmax = fmax(dmax, R1[j * IB + j]); }
            if (!(dmin > FAST_SPREAD_MIN * dmax)) bad = 1;
        }
        __syncthreads();
        if (!bad) {
            if (PANEL_VARIANT & 4) trsm_dmma(R1, rinv, 0); else trsm(R1);
            stamp(3);
            gram_exchange(1, ftag + 1);
            stamp(4);
            // CholeskyQR2 is as good as Householder QR iff the first pass left Q1 reasonably orthonormal:
            // ||Q1'Q1 - I|| <= 1/4 bounds kappa(Q1) by 1.3 and the second pass restores orthogonality to O(eps).
            for (int x = tid; x < IB * IB; x += PANEL_THREADS) {
                const int i = x / IB, j = x % IB;
                if (!(fabs(Gm[i * LDG + j] - (i == j ? 1.0 : 0.0)) <= 0.25 / IB)) bad = 1;   // max-norm test, scaled for the 2-norm
            }
            __syncthreads();
            if (!bad) chol(R2);
            stamp(5);
        }
        __syncthreads();
        if (!bad) {
            if (PANEL_VARIANT & 4) trsm_dmma(R2, rinv, 0); else trsm(R2);
            stamp(6);
            // Rt = R2 * R1 (upper) -> Gm
            for (int x = tid; x < IB * IB; x += PANEL_THREADS) {
                const int i = x / IB, k = x % IB;
                double sacc = 0.0;
                if (k >= i)
                    for (int j = i; j <= k; ++j) sacc += R2[i * IB + j] * R1[j * IB + k];
                Gm[i * LDG + k] = sacc;
            }
            __syncthreads();      // R2 is free from here: it becomes Up, the frozen rows U(j, j+1:) with aligned rows
            double* Up = R2;
            if (cta == 0) {
                // LU of the top block of E - Q S on CTA 0: Wt(i,k) = Q(i,k); row j is frozen at step j.  Block-wide with a
                // barrier per step: a one-warp register version (shuffles) was 2x slower, and pre-computing the next pivot
                // to take the division off the chain gained nothing (profiles/r01_panel_variants.txt).
                for (int x = tid; x < IB * IB; x += PANEL_THREADS) Wt[(x / IB) * LDG + (x % IB)] = S[(x % IB) * lds + (x / IB)];
                __syncthreads();
                for (int j = 0; j < IB; ++j) {
                    const double w = Wt[j * LDG + j];
                    const double sgn = w > 0.0 ? -1.0 : 1.0;
                    const double u = 1.0 + fabs(w);
                    if (tid == 0) { Sg[j] = sgn; Ud[j] = u; }
                    const double f = sgn / u;   // -l_i = f * W(i,j)
                    for (int x = tid; x < IB * IB; x += PANEL_THREADS) {
                        const int i = x / IB, k = x % IB;
                        if (i > j && k > j) Wt[i * LDG + k] += f * Wt[i * LDG + j] * Wt[j * LDG + k];
                    }
                    __syncthreads();
                }
                for (int x = tid; x < IB * IB; x += PANEL_THREADS)
                    if (x % IB > x / IB) { const double v = Wt[(x / IB) * LDG + (x % IB)]; ll_store(c2u(x), v, ftag + 2); Up[x] = v; }
                if (tid < IB) { ll_store(c2u(IB * IB + tid), Ud[tid], ftag + 2); ll_store(c2u(IB * IB + IB + tid), Sg[tid], ftag + 2); }
            } else {
                {   // all cells of a thread requested before the first wait
                    unsigned long long w0[4], w1[4];
                    const int xs[4] = {tid, tid + PANEL_THREADS, IB * IB + tid, IB * IB + IB + tid};
                    const bool on[4] = {tid % IB > tid / IB, (tid + PANEL_THREADS) % IB > (tid + PANEL_THREADS) / IB, tid < IB, tid < IB};
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (on[u]) ll_peek(c2u(xs[u]), w0[u], w1[u]);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (!on[u]) continue;
                        const double v = ll_finish(c2u(xs[u]), w0[u], w1[u], ftag + 2, a.backoff);
                        if (u < 2) Up[xs[u]] = v;
                        else if (u == 2) Ud[tid] = v;
                        else Sg[tid] = v;
                    }
                }
            }
            __syncthreads();
            stamp(7);
            if (tid < IB) { rsq[tid] = 1.0 / sqrt(Ud[tid]); cl[tid] = -Sg[tid] / Ud[tid]; }
            __syncthreads();
            // rows below the top block: L2 = M2 U^{-1}, scaled to the reference's |v|^2 = 2 convention
            if (PANEL_VARIANT & 4) {
                // the same recurrence as a triangular solve V = M Rr^{-1}, Rr = diag(sqrt(Ud)) (I + diag(cl) striu(U)), in R1
                for (int x = tid; x < IB * IB; x += PANEL_THREADS) {
                    const int i = x / IB, k = x % IB;
                    const double sq = Ud[i] * rsq[i];
                    R1[x] = k > i ? (cl[i] * Up[x]) * sq : (k == i ? sq : 0.0);
                }
                __syncthreads();
                trsm_dmma(R1, rsq, cta == 0 ? IB / 8 : 0);
            } else
            for (int r = tid; r < nr; r += PANEL_THREADS) {
                if (row0 + r < IB) continue;
                double x[IB];
#pragma unroll
                for (int k = 0; k < IB; ++k) x[k] = S[k * lds + r];
#pragma unroll
                for (int j = 0; j < IB; ++j) {
                    const double wj = x[j];
                    const double l = wj * cl[j];
                    x[j] = wj * rsq[j];
                    DHQR_ROW_AXPY(x, l, Up + j * IB, j)
                    asm volatile("" ::: "memory");
                }
#pragma unroll
                for (int k = 0; k < IB; ++k) S[k * lds + r] = x[k];
            }
            if (cta == 0) {   // top block: V below the diagonal, R above, alpha
                for (int x = tid; x < IB * IB; x += PANEL_THREADS) {
                    const int i = x / IB, j = x % IB;   // row i, column j
                    double v;
                    if (i > j) v = Wt[i * LDG + j] * rsq[j];
                    else if (i == j) v = -Sg[j] * (Ud[j] * rsq[j]);
                    else v = Sg[i] * Gm[i * LDG + j];
                    S[j * lds + i] = v;
                }
                if (tid < IB) a.alpha[tid] = Sg[tid] * Gm[tid * LDG + tid];
            }
            done = true;
            stamp(8);
        } else {
            // fall back: the slab was modified by the first TRSM at most; reload the untouched panel from memory
            __syncthreads();
            for (int c = warp; c < nc; c += PNW)
                for (int r = lane; r < nr4; r += 32) S[c * lds + r] = r < nr ? a.P[(int64_t)c * a.ldp + row0 + r] : 0.0;
        }
        __syncthreads();
        if (a.fast_stats && cta == 0 && tid == 0) atomicAdd(&a.fast_stats[done ? 0 : 1], 1);
    }

    if (!done) {
    // produce(jn): [apply reflector jn-1 to the columns right of jn]  +  partial dots of column jn (rows >= jn)
    // against the columns >= jn, published as cells of step jn; CTA 0 also publishes row jn (the next pivot row);
    // owner warps gather the partials of their column and publish the total.
    auto produce = [&](const int jn, const bool upd, const double f, const double alpha, const int pb) {
        const uint32_t tag = a.epoch + 1 + jn;
        const int jp = jn - 1;
        const int r_lo1 = clampi((int64_t)jn - row0);                 // first local row that enters the dots
        const int rstart = upd ? clampi((int64_t)jp - row0) : r_lo1;   // first local row touched by the update
        for (int cA = jn + warp; cA < nc; cA += 2 * PNW) {
            const int cB = cA + PNW;
            const bool hasB = cB < nc;
            const bool updA = upd && cA != jn;                          // column jn itself was updated in step 1
            const double wA = updA ? f * (tot[pb][cA] - alpha * pv[pb][cA]) : 0.0;
            const double wB = (upd && hasB) ? f * (tot[pb][cB] - alpha * pv[pb][cB]) : 0.0;
            double accA = 0.0, accB = 0.0, pivA = 0.0, pivB = 0.0;
            const double* xcol = S + jn * lds;
            const double* vcol = S + (upd ? jp : jn) * lds;
            double* colA = S + cA * lds;
            double* colB = S + (hasB ? cB : cA) * lds;
            double accA2 = 0.0, accB2 = 0.0;
            for (int r = rstart + lane; r < nr; r += 64) {   // two independent rows per iteration (ILP)
                const int r2 = r + 32;
                const bool ok2 = r2 < nr;
                const double xn = xcol[r], xn2 = ok2 ? xcol[r2] : 0.0;
                const double v = upd ? vcol[r] : 0.0, v2 = (upd && ok2) ? vcol[r2] : 0.0;
                double tA = colA[r], tB = hasB ? colB[r] : 0.0;
                double tA2 = ok2 ? colA[r2] : 0.0, tB2 = (ok2 && hasB) ? colB[r2] : 0.0;
                if (updA) { tA -= v * wA; colA[r] = tA; if (ok2) { tA2 -= v2 * wA; colA[r2] = tA2; } }
                if (upd && hasB) { tB -= v * wB; colB[r] = tB; if (ok2) { tB2 -= v2 * wB; colB[r2] = tB2; } }
                if (r >= r_lo1) { accA += xn * tA; accB += xn * tB; }
                if (ok2 && r2 >= r_lo1) { accA2 += xn2 * tA2; accB2 += xn2 * tB2; }
                if (row0 + r == jn) { pivA = tA; pivB = tB; }
                if (ok2 && row0 + r2 == jn) { pivA = tA2; pivB = tB2; }
            }
            accA += accA2;
            accB += accB2;
            accA = warp_sum(accA);
            accB = warp_sum(accB);
            if (lane == 0) {
                ll_store(pcell(jn, cta, cA), accA, tag);
                if (hasB) ll_store(pcell(jn, cta, cB), accB, tag);
            }
            if (cta == 0) {   // exactly one lane holds row jn of this slab
                const int64_t d = (int64_t)jn - rstart;
                if (d >= 0 && d < nr && (d & 31) == lane) {
                    ll_store(vcell(jn, cA), pivA, tag);
                    if (hasB) ll_store(vcell(jn, cB), pivB, tag);
                }
            }
            // owner gather (fixed order: lane l sums CTAs l, l+32, ...; then the shuffle tree)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = h ? cB : cA;
                if (a.levels != 2 || (h && !hasB) || (c % G) != cta) continue;
                unsigned long long w0[PANEL_MAXG / 32], w1[PANEL_MAXG / 32];
#pragma unroll
                for (int t = 0; t < PANEL_MAXG / 32; ++t)
                    if (lane + 32 * t < G) ll_peek(pcell(jn, lane + 32 * t, c), w0[t], w1[t]);
                double sum = 0.0;
#pragma unroll
                for (int t = 0; t < PANEL_MAXG / 32; ++t)
                    if (lane + 32 * t < G) sum += ll_finish(pcell(jn, lane + 32 * t, c), w0[t], w1[t], tag, a.backoff);
                sum = warp_sum(sum);
                if (lane == 0) ll_store(tcell(jn, c), sum, tag);
            }
        }
    };

    long long* tr = a.trace ? a.trace + (size_t)cta * IB * 8 : nullptr;
    const long long tstart = clock64();
    produce(0, false, 0.0, 0.0, 0);

    for (int j = 0; j < nc; ++j) {
        const uint32_t tag = a.epoch + 1 + j;
        const int pb = j & 1;
        if (tr && tid == 0) tr[j * 8 + 0] = clock64() - tstart;   // enter iteration
        if (a.levels == 2) {
            if (warp == 0) {
                if (lane >= j && lane < nc) tot[pb][lane] = ll_wait(tcell(j, lane), tag, a.backoff);
                if (tr && lane == j) tr[j * 8 + 7] = clock64() - tstart;   // total of column j arrived
            } else if (warp == 1) {
                if (lane >= j && lane < nc) pv[pb][lane] = ll_wait(vcell(j, lane), tag, a.backoff);
                if (tr && lane == j) tr[j * 8 + 6] = clock64() - tstart;   // pivot element arrived
            }
        } else {
            // one hand-off: every CTA sums all G partials of every live column itself (same fixed order as the
            // owner gather: lane l takes CTAs l, l+32, ...; then the shuffle tree) -> results identical on all CTAs
            for (int c = j + warp; c < nc; c += PNW) {
                unsigned long long w0[PANEL_MAXG / 32], w1[PANEL_MAXG / 32];
#pragma unroll
                for (int t = 0; t < PANEL_MAXG / 32; ++t)
                    if (lane + 32 * t < G) ll_peek(pcell(j, lane + 32 * t, c), w0[t], w1[t]);
                double sum = 0.0;
#pragma unroll
                for (int t = 0; t < PANEL_MAXG / 32; ++t)
                    if (lane + 32 * t < G) sum += ll_finish(pcell(j, lane + 32 * t, c), w0[t], w1[t], tag, a.backoff);
                sum = warp_sum(sum);
                if (lane == 0) tot[pb][c] = sum;
                if (tr && c == j && lane == 0) tr[j * 8 + 7] = clock64() - tstart;
            }
            if (warp == PNW - 1 && lane >= j && lane < nc) pv[pb][lane] = ll_wait(vcell(j, lane), tag, a.backoff);
        }
        if (tr && tid == 0) tr[j * 8 + 1] = clock64() - tstart;   // warp 0 has its totals
        __syncthreads();
        if (tr && tid == 0) tr[j * 8 + 2] = clock64() - tstart;   // block has totals + pivot row
        // S:129-131
        const double xj = pv[pb][j];
        const double s = sqrt(tot[pb][j]);
        const double sg = xj > 0.0 ? 1.0 : (xj < 0.0 ? -1.0 : 0.0);   // sign(0) == 0 as in S:8
        const double alpha = -sg * s;
        const double f = 1.0 / sqrt(s * (s + fabs(xj)));
        if (cta == 0 && tid == 0) a.alpha[j] = alpha;
        if (tr && tid == 0) tr[j * 8 + 3] = clock64() - tstart + (long long)(f == 12345.678);   // scalars done
        // step 1: v = f (x - alpha e_j) in place; next pivot column updated in place (rows >= j)
        const bool has_next = j + 1 < nc;
        const double w1n = has_next ? f * (tot[pb][j + 1] - alpha * pv[pb][j + 1]) : 0.0;
        for (int r = clampi((int64_t)j - row0) + tid; r < nr; r += PANEL_THREADS) {
            double x = S[j * lds + r];
            if (row0 + r == j) x -= alpha;
            const double v = f * x;
            S[j * lds + r] = v;
            if (has_next) S[(j + 1) * lds + r] -= v * w1n;
        }
        __syncthreads();
        if (tr && tid == 0) tr[j * 8 + 4] = clock64() - tstart;   // step 1 done
        if (has_next) produce(j + 1, true, f, alpha, pb);
        if (tr && tid == 0) tr[j * 8 + 5] = clock64() - tstart;   // warp 0 finished produce (+ owner gather if any)
    }
    }   // if (!done)
    __syncthreads();
    // write back the factored slab, and the packed V block
    for (int c = warp; c < nc; c += PNW)
        for (int r = lane; r < nr; r += 32) a.P[(int64_t)c * a.ldp + row0 + r] = S[c * lds + r];
    if (a.vpk) {
        for (int c = warp; c < IB; c += PNW) {
            const int pc = a.voff + c;
            for (int r = lane; r < nr; r += 32)
                a.vpk[vpk_index(a.vtop + row0 + r, pc)] = (c < nc && row0 + r >= c) ? S[c * lds + r] : 0.0;
            if (cta == 0)
                for (int64_t r = lane; r < a.vtop; r += 32) a.vpk[vpk_index(r, pc)] = 0.0;
            if (cta == G - 1)
                for (int64_t r = a.vtop + a.mp + lane; r < a.vrows; r += 32) a.vpk[vpk_index(r, pc)] = 0.0;
        }
    }
    if (a.trace && tid == 0) a.trace[(size_t)cta * IB * 8 + 9] = clock64();   // absolute, for the write-back duration see [10]
}

// ------------------------------------------------------------------------------------------------
// pack: copy a Householder block into the packed V layout (vpk).  tril != 0: the block is stored in
// place in A (lower trapezoid incl. diagonal, S:232-242 reads H in place; the GEMM path wants zeros
// above the diagonal); tril == 0: plain copy (kernel-level test hook).  Columns >= kb and rows
// outside [vtop, vtop+mp) are zero-filled.  grid.y = packed columns to write.
// ------------------------------------------------------------------------------------------------
__global__ void k_pack(const double* __restrict__ A, int64_t lda, int64_t mp, int kb, int tril, double* __restrict__ vpk,
                       int voff, int64_t vtop, int64_t vrows) {
    // four consecutive window rows per thread (they never straddle a 64-row chunk): two 16-byte stores, and two 16-byte loads
    // when the source is aligned
    const int c = blockIdx.y;
    const bool fast = ((vtop & 3) == 0) && ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    for (int64_t r = 4 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x); r < vrows; r += 4 * (int64_t)gridDim.x * blockDim.x) {
        const int64_t pr = r - vtop;   // panel-relative row of the first of the four
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        if (c < kb) {
            const double* src = A + (int64_t)c * lda + pr;
            if (fast && pr >= 0 && pr + 3 < mp) {
                const double2 a = *reinterpret_cast<const double2*>(src), b = *reinterpret_cast<const double2*>(src + 2);
                v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (pr + q >= 0 && pr + q < mp) v[q] = src[q];
            }
            if (tril) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (pr + q < c) v[q] = 0.0;
            }
        }
        double* dst = vpk + vpk_index(r, voff + c);
        *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2*>(dst + 2) = make_double2(v[2], v[3]);
    }
}
// ------------------------------------------------------------------------------------------------
// Q'b / Qb with ONE right-hand side (S:226-242).  The sweep over the panels is sequential, but once T' of every panel is
// known (b-independent: packed Gram matrices + a batched k_tinv before the sweep) a panel costs two GEMV-shaped passes over
// its reflectors, read IN PLACE from the factored matrix (lower trapezoid including the diagonal, S:232-242):
//   k_qt_dot : w = V'b, one partial per CTA (16 independent column accumulators per lane = S:42-49 for 16 columns at once,
//              warp-shuffle reduction); the CTA that arrives last sums the partials in CTA order (deterministic) and forms
//              y = -T'w (Q'b) or y = -Tw (Qb)
//   k_qt_axpy: b += V y   (S:156-160 for the whole panel; V comes from L2, the first pass just read it)
// ------------------------------------------------------------------------------------------------
constexpr int QT_THREADS = 512;            // k_qt_dot: 16 warps x 8 columns, one CTA per SM
constexpr int QT_ATHREADS = 256;           // k_qt_axpy: 64 rows x 4 column quarters
constexpr int QT_AROWS = 64;
constexpr int QT_MAXROWS = 1024;           // rows of b a k_qt_dot CTA keeps in shared memory
struct QtArgs {
    const double* V;        // first column of the panel at its pivot row
    int64_t lda;
    int64_t mp;             // rows from the pivot row to the end
    int kb;                 // reflectors in the panel (<= 128)
    double* b;              // right-hand side at the pivot row
    const double* Linv;     // T' of the panel: 128 x 128, element (i, k) at k * 128 + i, lower triangular
    double* part;           // [gridDim.x][128]
    double* y;              // [128]
    unsigned int* ticket;   // zero on entry, zero again on exit
    int rows_per_cta;       // multiple of 32, <= QT_MAXROWS
    int trans;
};

__global__ void __launch_bounds__(QT_THREADS, 1) k_qt_dot(QtArgs a) {
    __shared__ double sb[QT_MAXROWS];
    __shared__ double sw[4][WP];
    __shared__ int s_last;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t r0 = (int64_t)blockIdx.x * a.rows_per_cta;
    const int nr = (int)max((int64_t)0, min((int64_t)a.rows_per_cta, a.mp - r0));
    for (int r = tid; r < nr; r += QT_THREADS) sb[r] = a.b[r0 + r];
    __syncthreads();
    // warp w owns columns 8 w .. 8 w + 7; lane: rows r0 + lane + 32 i
    const int c0 = warp * 8;
    double acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0;
    if (c0 < a.kb) {
        const double* v = a.V + (int64_t)c0 * a.lda + r0;
        if ((c0 + 8 <= a.kb) && (r0 >= c0 + 7)) {                    // every column live, every row below the diagonal
            int r = lane;
            for (; r + 32 < nr; r += 64) {                            // 16 independent loads in flight per lane
                const double b0 = sb[r], b1 = sb[r + 32];
                double x[8], z[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { x[j] = v[(int64_t)j * a.lda + r]; z[j] = v[(int64_t)j * a.lda + r + 32]; }
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += x[j] * b0 + z[j] * b1;
            }
            for (; r < nr; r += 32) {
                const double b0 = sb[r];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += v[(int64_t)j * a.lda + r] * b0;
            }
        } else {
            for (int r = lane; r < nr; r += 32) {
                const double b0 = sb[r];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (c0 + j < a.kb && r0 + r >= c0 + j) acc[j] += v[(int64_t)j * a.lda + r] * b0;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = warp_sum(acc[j]);
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a.part[(int64_t)blockIdx.x * WP + c0 + j] = acc[j];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // the CTA that arrived last: w = sum of the partials in CTA order (four contiguous slices, then a fixed tree), y = -T'w
    const int c = tid & (WP - 1), sl = tid >> 7;
    {
        const int G = (int)gridDim.x, gq = (G + 3) / 4, g0 = sl * gq, g1 = min(G, g0 + gq);
        double s = 0.0;
        for (int g = g0; g < g1; g += 12) {                          // twelve independent loads in flight, added in CTA order
            double v[12];
#pragma unroll
            for (int u = 0; u < 12; ++u) v[u] = (g + u < g1) ? __ldcg(&a.part[(int64_t)(g + u) * WP + c]) : 0.0;
#pragma unroll
            for (int u = 0; u < 12; ++u) s += v[u];
        }
        sw[sl][c] = s;
    }
    __syncthreads();
    if (tid < WP) sb[tid] = (sw[0][tid] + sw[1][tid]) + (sw[2][tid] + sw[3][tid]);
    __syncthreads();
    {
        const int i = c, k0 = sl * 32;
        double s0 = 0.0, s1 = 0.0;
        if (!a.trans) {                                               // y = -T' w: row i of the lower triangular T'
#pragma unroll 8
            for (int k = k0; k < k0 + 32; k += 2) {
                if (k <= i) s0 += a.Linv[k * WP + i] * sb[k];
                if (k + 1 <= i) s1 += a.Linv[(k + 1) * WP + i] * sb[k + 1];
            }
        } else {                                                      // y = -T w: column i of T'
#pragma unroll 8
            for (int k = k0; k < k0 + 32; k += 2) {
                if (k >= i) s0 += a.Linv[i * WP + k] * sb[k];
                if (k + 1 >= i) s1 += a.Linv[i * WP + k + 1] * sb[k + 1];
            }
        }
        sw[sl][i] = s0 + s1;
    }
    __syncthreads();
    if (tid < WP) a.y[tid] = -((sw[0][tid] + sw[1][tid]) + (sw[2][tid] + sw[3][tid]));
    if (tid == 0) *a.ticket = 0u;
}

__global__ void __launch_bounds__(QT_ATHREADS) k_qt_axpy(QtArgs a) {
    __shared__ double sy[WP];
    __shared__ double sp[4][QT_AROWS];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < WP) sy[tid] = a.y[tid];
    __syncthreads();
    const int rl = (warp & 1) * 32 + lane, q = warp >> 1;            // row within the CTA, column quarter
    const int64_t r = (int64_t)blockIdx.x * QT_AROWS + rl;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (r < a.mp) {
        const int c0 = q * 32, c1 = (int)min((int64_t)min(a.kb, c0 + 32), r + 1);   // columns with row r at or below their diagonal
        const double* v = a.V + r;
        int c = c0;
        for (; c + 8 <= c1; c += 8) {
            const double v0 = v[(int64_t)c * a.lda], v1 = v[(int64_t)(c + 1) * a.lda], v2 = v[(int64_t)(c + 2) * a.lda], v3 = v[(int64_t)(c + 3) * a.lda];
            const double v4 = v[(int64_t)(c + 4) * a.lda], v5 = v[(int64_t)(c + 5) * a.lda], v6 = v[(int64_t)(c + 6) * a.lda], v7 = v[(int64_t)(c + 7) * a.lda];
            s0 += v0 * sy[c] + v4 * sy[c + 4];
            s1 += v1 * sy[c + 1] + v5 * sy[c + 5];
            s2 += v2 * sy[c + 2] + v6 * sy[c + 6];
            s3 += v3 * sy[c + 3] + v7 * sy[c + 7];
        }
        for (; c < c1; ++c) s0 += v[(int64_t)c * a.lda] * sy[c];
    }
    sp[q][rl] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (tid < QT_AROWS) {
        const int64_t rr = (int64_t)blockIdx.x * QT_AROWS + tid;
        if (rr < a.mp) a.b[rr] += (sp[0][tid] + sp[1][tid]) + (sp[2][tid] + sp[3][tid]);
    }
}

// zero packed columns [c0, c1) over all chunks
__global__ void k_vpk_zero_cols(double* __restrict__ vpk, int64_t nchunks, int c0, int c1) {
    const int64_t per = (int64_t)(c1 - c0) * LD1;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nchunks * per; e += (int64_t)gridDim.x * blockDim.x)
        vpk[(e / per) * VPK_CHUNK + (int64_t)c0 * LD1 + e % per] = 0.0;
}

// ------------------------------------------------------------------------------------------------
// back-substitution step (S:256-282), column oriented: solve the bs x bs diagonal block
//   x_blk = R_bb^{-1} y_blk   (R_bb = triu(A_bb,1) + diag(alpha))   in every CTA (one warp),
// then y[0:c0) -= R[0:c0, blk] x_blk on the CTA's slice of rows.  CTA 0 publishes x_blk.
// ------------------------------------------------------------------------------------------------
constexpr int BS_BLK = 32;
__global__ void __launch_bounds__(256) k_backsolve_step(const double* __restrict__ Ablk, int64_t lda,
                                                        const double* __restrict__ alpha, double* __restrict__ y,
                                                        int64_t ldy, int nrhs, double* __restrict__ x, int64_t ldx,
                                                        int64_t c0, int bs) {
    // Ablk: pointer to (row 0, first column of the block) in local storage; c0 = global index of that column
    __shared__ double sx[BS_BLK];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int rhs = 0; rhs < nrhs; ++rhs) {
        double* yr = y + (int64_t)rhs * ldy;
        if (tid < 32) {
            double yk = lane < bs ? yr[c0 + lane] : 0.0;
            for (int i = bs - 1; i >= 0; --i) {
                const double xi = __shfl_sync(0xffffffffu, yk, i) / alpha[c0 + i];
                if (lane == i) yk = xi;
                if (lane < i) yk -= Ablk[(int64_t)i * lda + c0 + lane] * xi;
            }
            sx[lane] = yk;
        }
        __syncthreads();
        if (blockIdx.x == 0 && tid < bs) x[(int64_t)rhs * ldx + c0 + tid] = sx[tid];
        for (int64_t r = (int64_t)blockIdx.x * blockDim.x + tid; r < c0; r += (int64_t)gridDim.x * blockDim.x) {
            double acc = 0.0;
            for (int k = 0; k < bs; ++k) acc += Ablk[(int64_t)k * lda + r] * sx[k];
            yr[r] -= acc;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// back-substitution as ONE launch (S:256-282, column oriented): a wavefront over 32-row strips.
//   CTA (NLOW + k) owns the diagonal strip of local block k (rows = columns [col0 + 32k, +bs)): it keeps its piece of y in
//   shared memory, subtracts R[strip, block b] x_b for the later blocks b = last .. k+1 as their x_b appear, then solves its own
//   32 x 32 diagonal block (warp-shuffle substitution, diag(R) = alpha) and publishes x_k.  CTAs 0 .. NLOW-1 own the rows above
//   this rank's columns (strips of 32 rows of [0, col0)): update only.  x_b travels in self-validating cells (value + launch
//   tag in one 16-byte store, like the panel kernel's exchange): one L2 round trip from "solved" to "seen", no flags, no
//   fences; the tile of R a CTA needs next is loaded BEFORE it starts polling, so the critical path per block is
//   substitution + one L2 hand-off + a 32 x 32 matrix-vector product.  All CTAs must be co-resident (they spin): the host checks.
// ------------------------------------------------------------------------------------------------
constexpr int BW_THREADS = 128;
__global__ void __launch_bounds__(BW_THREADS) k_backsolve_wave(const double* __restrict__ A, int64_t lda, const double* __restrict__ alpha,
                                                               double* __restrict__ y, double* __restrict__ x, int64_t col0, int64_t nl,
                                                               int nlow, unsigned long long* cells, uint32_t tag) {
    __shared__ double sy[32], sx[32], part[4][32], sR[32][33];
    const int tid = threadIdx.x, lane = tid & 31, grp = tid >> 5;
    const int nbk = (int)((nl + 31) / 32);
    const bool diag = (int)blockIdx.x >= nlow;
    const int k = diag ? (int)blockIdx.x - nlow : -1;                       // own block (diag strips)
    const int64_t r0 = diag ? col0 + 32 * (int64_t)k : 32 * (int64_t)blockIdx.x;   // first row of the strip
    const int nr = (int)min((int64_t)32, (diag ? col0 + nl : col0) - r0);          // rows in the strip
    if (tid < 32) sy[tid] = tid < nr ? y[r0 + tid] : 0.0;
    if (diag) {                                                                // own diagonal block: triu(A_bb, 1), by columns
        for (int e = tid; e < 32 * 32; e += BW_THREADS) {
            const int r = e & 31, cc = e >> 5;
            sR[r][cc] = (r < cc && cc < nr) ? A[(32 * (int64_t)k + cc) * lda + r0 + r] : 0.0;
        }
    }
    __syncthreads();
    for (int b = nbk - 1; b > k; --b) {
        const int bs = (int)min((int64_t)32, nl - 32 * (int64_t)b);
        // this thread's 8 entries of R[strip, block b]: row `lane`, columns 8 grp .. 8 grp + 7 (issued before the wait)
        double rv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int cc = 8 * grp + q;
            rv[q] = (lane < nr && cc < bs) ? A[(32 * (int64_t)b + cc) * lda + r0 + lane] : 0.0;
        }
        if (tid < 32) sx[tid] = tid < bs ? ll_wait(cells + ((size_t)b * 32 + tid) * 2, tag, 0) : 0.0;
        __syncthreads();
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) acc += rv[q] * sx[8 * grp + q];
        part[grp][lane] = acc;
        __syncthreads();
        if (tid < 32) sy[tid] -= (part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]);
        __syncthreads();
    }
    if (!diag) {
        if (tid < nr) y[r0 + tid] = sy[tid];
        return;
    }
    if (tid < 32) {                                                            // x_k = R_kk^{-1} y_k (S:266-267, i = last .. first)
        double yk = sy[lane];
        for (int i = nr - 1; i >= 0; --i) {
            const double xi = __shfl_sync(0xffffffffu, yk, i) / alpha[r0 + i];
            if (lane == i) yk = xi;
            if (lane < i) yk -= sR[lane][i] * xi;
        }
        if (lane < nr) {
            ll_store(cells + ((size_t)k * 32 + lane) * 2, yk, tag);
            x[r0 + lane] = yk;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// unblocked path (nb = 1, BASELINE config 2): one reflector per step.
//   k_house1: S:129-135 for column j (one CTA): norm via warp-shuffle tree, scale in place, and a
//             16B-aligned copy of v (zero-padded to a multiple of 2) for TMA staging.
//   k_apply1: S:198-213: each CTA owns CW trailing columns; v and the column tile are staged into
//             shared memory with TMA bulk copies, one warp-shuffle dot + axpy per column, written
//             back once (one read + one write of the trailing matrix per step).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1) k_house1(double* __restrict__ col, int64_t len, double* __restrict__ alpha,
                                                    double* __restrict__ vout) {
    __shared__ double red[32];
    __shared__ double sc[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double acc = 0.0;
    for (int64_t i = tid; i < len; i += 1024) {
        const double x = col[i];
        acc += x * x;
    }
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (warp == 0) {
        double t = red[lane];
        t = warp_sum(t);
        if (lane == 0) {
            const double x0 = col[0];
            const double s = sqrt(t);
            const double sg = x0 > 0.0 ? 1.0 : (x0 < 0.0 ? -1.0 : 0.0);
            const double al = -sg * s;
            *alpha = al;
            sc[0] = al;
            sc[1] = 1.0 / sqrt(s * (s + fabs(x0)));
        }
    }
    __syncthreads();
    const double al = sc[0], f = sc[1];
    for (int64_t i = tid; i < len; i += 1024) {
        double x = col[i];
        if (i == 0) x -= al;
        x *= f;
        col[i] = x;
        vout[i] = x;
    }
    if (tid == 0) vout[len] = 0.0;   // pad so the staged copy is a whole number of 16B units
}

constexpr int A1_CW = 2;         // columns per CTA
constexpr int A1_THREADS = 256;
// staged variant: requires (len_pad * 8 * (A1_CW + 1)) bytes of smem, 16B-aligned column starts.
// Fused next reflector (vnext != null, single GPU): the first column of CTA 0 is column j+1; once it is updated the same CTA
// runs S:129-135 on it (norm, alpha, scale: what k_house1 does) and leaves v_{j+1} in vnext for the next launch, so a column
// step is ONE launch instead of two and the one-CTA k_house1 leaves the critical path.  `lead` = j & 1 (the window starts on
// an even row); the next window starts at row j + 1 - lead_next with lead_next = 1 - lead.
__global__ void __launch_bounds__(A1_THREADS) k_apply1_tma(const double* __restrict__ v, int64_t len,
                                                          double* __restrict__ C, int64_t ldc, int ncols, int aligned,
                                                          double* __restrict__ vnext, double* __restrict__ alpha_next, int lead) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ double red[A1_CW][A1_THREADS / 32];
    __shared__ double hs[2];
    const int64_t lenp = (len + 1) & ~(int64_t)1;
    double* sv = reinterpret_cast<double*>(smem_raw);
    double* sc = sv + lenp;   // [A1_CW][lenp]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c0 = blockIdx.x * A1_CW;
    const int nc = min(A1_CW, ncols - c0);
    const int64_t nb = aligned ? (len & ~(int64_t)1) : 0;   // elements per column moved by TMA
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    // Programmatic dependent launch: column step j+1 is launched while step j still runs, so that its CTAs are resident and
    // past their prologue the moment step j's memory is complete (every byte they read was written by step j).
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();
    // generic fill of what TMA cannot move
    for (int c = 0; c < nc; ++c)
        for (int64_t i = nb + tid; i < len; i += A1_THREADS) sc[c * lenp + i] = C[(int64_t)(c0 + c) * ldc + i];
    __syncthreads();
    if (tid == 0) {
        uint32_t bytes = (uint32_t)(lenp * 8) + (uint32_t)(nc * nb * 8);
        mbar_arrive_expect_tx(&bar, bytes);
        // chunked: a single bulk copy is limited in size; 32 KB pieces
        for (int64_t o = 0; o < lenp; o += 4096) bulk_g2s(sv + o, v + o, (uint32_t)(min((int64_t)4096, lenp - o) * 8), &bar);
        for (int c = 0; c < nc; ++c)
            for (int64_t o = 0; o < nb; o += 4096)
                bulk_g2s(sc + c * lenp + o, C + (int64_t)(c0 + c) * ldc + o, (uint32_t)(min((int64_t)4096, nb - o) * 8), &bar);
    }
    mbar_wait(&bar, 0);
    for (int c = 0; c < nc; ++c) {
        double acc = 0.0;
        for (int64_t i = tid; i < len; i += A1_THREADS) acc += sv[i] * sc[c * lenp + i];   // S:208 partialdot
        acc = warp_sum(acc);
        if (lane == 0) red[c][warp] = acc;
    }
    __syncthreads();
    const bool next = vnext != nullptr && blockIdx.x == 0;
    for (int c = 0; c < nc; ++c) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < A1_THREADS / 32; ++w) s += red[c][w];
        double* out = C + (int64_t)(c0 + c) * ldc;
        for (int64_t i = tid; i < len; i += A1_THREADS) {
            const double x = sc[c * lenp + i] - sv[i] * s;                                   // S:209 hotloop!
            out[i] = x;
            if (next && c == 0) sc[i] = x;
        }
    }
    if (!next) return;
    // S:129-135 for column j+1 (rows >= j+1 = window rows >= i1)
    const int64_t i1 = 1 + lead;
    __syncthreads();
    double acc = 0.0;
    for (int64_t i = i1 + tid; i < len; i += A1_THREADS) acc += sc[i] * sc[i];
    acc = warp_sum(acc);
    if (lane == 0) red[0][warp] = acc;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < A1_THREADS / 32; ++w) t += red[0][w];
        const double x0 = sc[i1];
        const double sn = sqrt(t);
        const double sg = x0 > 0.0 ? 1.0 : (x0 < 0.0 ? -1.0 : 0.0);
        const double al = -sg * sn;
        *alpha_next = al;
        hs[0] = al;
        hs[1] = 1.0 / sqrt(sn * (sn + fabs(x0)));
    }
    __syncthreads();
    const double al = hs[0], f = hs[1];
    const int leadn = 1 - lead;
    for (int64_t i = i1 + tid; i < len; i += A1_THREADS) {
        double x = sc[i];
        if (i == i1) x -= al;
        x *= f;
        C[i] = x;                                   // column j+1 of the matrix: the reflector in place (S:133-135)
        vnext[leadn + (i - i1)] = x;
    }
    if (tid == 0) {
        if (leadn) vnext[0] = 0.0;
        vnext[leadn + (len - i1)] = 0.0;            // pad: the staged copy moves whole 16-byte units
    }
}
// direct variant (column tile does not fit in shared memory): two passes, second read hits L2
__global__ void __launch_bounds__(A1_THREADS) k_apply1_direct(const double* __restrict__ v, int64_t len,
                                                             double* __restrict__ C, int64_t ldc, int ncols) {
    __shared__ double red[A1_THREADS / 32];
    __shared__ double sdot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int c = blockIdx.x; c < ncols; c += gridDim.x) {
        double* col = C + (int64_t)c * ldc;
        double acc = 0.0;
        for (int64_t i = tid; i < len; i += A1_THREADS) acc += v[i] * col[i];
        acc = warp_sum(acc);
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < A1_THREADS / 32; ++w) s += red[w];
            sdot = s;
        }
        __syncthreads();
        const double s = sdot;
        for (int64_t i = tid; i < len; i += A1_THREADS) col[i] -= v[i] * s;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// unblocked path as ONE persistent launch (single GPU, m <= UW_MAXI * UW_THREADS rows): the whole column loop S:127-144.
//   CTA g owns the columns c == g (mod G) for the whole factorisation.  Step j: every CTA waits for "v_j is in place"
//   (a release/acquire flag per column), reads v_j = A[j:m, j] into registers, and for each of its columns c > j reads the
//   column tail into registers, forms s = v'a (S:208, warp shuffles + one block reduction) and writes back a - v s (S:209).
//   The CTA that owns column j+1 takes it first and runs S:129-135 on it straight away (norm, alpha, scale: the next
//   reflector), publishes it and only then turns to its other columns: the reflector chain never waits for the trailing update.
//   No grid barrier: columns are private to their CTA, the only cross-CTA dependency is the reflector itself.
//   Everything streams L2 <-> registers (the 64 MiB matrix of BASELINE config 2 is L2 resident); CTAs spin on the flags, so all
//   of them must be resident: cooperative launch, G <= #SMs.
// ------------------------------------------------------------------------------------------------
constexpr int UW_THREADS = 256;
constexpr int UW_MAXI = 32;            // rows per thread: m <= 8192
__global__ void __launch_bounds__(UW_THREADS, 1) k_unblocked_wave(double* __restrict__ A, int64_t lda, int64_t m, int n,
                                                                  double* __restrict__ alpha, unsigned int* flags, unsigned int tag) {
    __shared__ double red[2][UW_THREADS / 32];
    __shared__ double hs[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, g = blockIdx.x;
    int slot = 0;
    auto block_sum = [&](double v) {             // fixed order: deterministic
        v = warp_sum(v);
        if (lane == 0) red[slot][warp] = v;
        __syncthreads();
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < UW_THREADS / 32; ++w) t += red[slot][w];
        slot ^= 1;
        return t;
    };
    // S:129-135 on the column tail x[] (rows r0 + tid + UW_THREADS i, i.e. rows >= r0) held in registers; writes alpha[jn]
    auto house = [&](double (&x)[UW_MAXI], int64_t r0, int jn) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < UW_MAXI; ++i) {
            const int64_t r = r0 + tid + (int64_t)UW_THREADS * i;
            if (r >= jn && r < m) acc += x[i] * x[i];
        }
        const double t = block_sum(acc);
        if (tid == (int)(jn - r0)) {             // the thread that holds row jn (jn - r0 is 0 or 1)
            const double x0 = x[0];
            const double sn = sqrt(t);
            const double sg = x0 > 0.0 ? 1.0 : (x0 < 0.0 ? -1.0 : 0.0);
            const double al = -sg * sn;
            alpha[jn] = al;
            hs[0] = al;
            hs[1] = 1.0 / sqrt(sn * (sn + fabs(x0)));
        }
        __syncthreads();
        const double al = hs[0], f = hs[1];
#pragma unroll
        for (int i = 0; i < UW_MAXI; ++i) {
            const int64_t r = r0 + tid + (int64_t)UW_THREADS * i;
            if (r >= jn && r < m) x[i] = (r == jn ? x[i] - al : x[i]) * f;
        }
    };
    auto publish = [&](int j) {
        __threadfence();
        __syncthreads();
        if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + j), "r"(tag) : "memory");
    };
    if (g == 0) {                                // column 0: nothing to apply first
        double x[UW_MAXI];
#pragma unroll
        for (int i = 0; i < UW_MAXI; ++i) {
            const int64_t r = tid + (int64_t)UW_THREADS * i;
            x[i] = r < m ? A[r] : 0.0;
        }
        house(x, 0, 0);
#pragma unroll
        for (int i = 0; i < UW_MAXI; ++i) {
            const int64_t r = tid + (int64_t)UW_THREADS * i;
            if (r < m) A[r] = x[i];
        }
        publish(0);
    }
    for (int j = 0; j + 1 < n; ++j) {
        int c = j + 1 + ((g - (j + 1)) % G + G) % G;        // this CTA's first column right of j
        if (c >= n) break;                                   // nothing left for this CTA: its later steps are empty as well
        double x[UW_MAXI];
        if (c == j + 1) {                                    // the next pivot column is ours: fetch it while v_j is still on its way
            __syncthreads();                                 // (the rows were written by other threads of this CTA in the step before)
            const double* col = A + (int64_t)c * lda;
#pragma unroll
            for (int i = 0; i < UW_MAXI; ++i) {
                const int64_t r = j + tid + (int64_t)UW_THREADS * i;
                x[i] = r < m ? __ldcg(col + r) : 0.0;
            }
        }
        if (tid == 0) {
            unsigned int f;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(flags + j) : "memory");
            } while (f != tag);
        }
        __syncthreads();
        double v[UW_MAXI];
        const double* vj = A + (int64_t)j * lda;
#pragma unroll
        for (int i = 0; i < UW_MAXI; ++i) {
            const int64_t r = j + tid + (int64_t)UW_THREADS * i;
            v[i] = r < m ? __ldcg(vj + r) : 0.0;      // written by another SM a moment ago: read it from L2, never from this SM's L1
        }
        for (; c < n; c += G) {
            double* col = A + (int64_t)c * lda;
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < UW_MAXI; ++i) {
                const int64_t r = j + tid + (int64_t)UW_THREADS * i;
                if (c != j + 1) x[i] = r < m ? __ldcg(col + r) : 0.0;
                acc += v[i] * x[i];                                                   // S:208 partialdot
            }
            const double s = block_sum(acc);
#pragma unroll
            for (int i = 0; i < UW_MAXI; ++i) x[i] -= v[i] * s;                       // S:209 hotloop!
            if (c == j + 1) house(x, j, j + 1);                                       // the next reflector, at once
#pragma unroll
            for (int i = 0; i < UW_MAXI; ++i) {
                const int64_t r = j + tid + (int64_t)UW_THREADS * i;
                if (r < m) col[r] = x[i];
            }
            if (c == j + 1) publish(j + 1);
        }
    }
}

// partialdot (S:42-49) as a standalone primitive: one CTA, warp-shuffle tree.
__global__ void __launch_bounds__(1024, 1) k_partialdot(const double* __restrict__ x, const double* __restrict__ y,
                                                        int64_t i0, int64_t i1, double* __restrict__ out) {
    __shared__ double red[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double acc = 0.0;
    for (int64_t i = i0 + tid; i < i1; i += 1024) acc += x[i] * y[i];
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (warp == 0) {
        double t = red[lane];
        t = warp_sum(t);
        if (lane == 0) *out = t;
    }
}

// counter-based U[0,1) fill, bit-identical to oracle/dhqr_oracle.c:dhqr_oracle_uniform
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__global__ void k_fill_uniform(uint64_t seed, int64_t i0, int64_t j0, int64_t m, int64_t n, double* __restrict__ A,
                               int64_t lda) {
    const uint64_t sh = mix64(seed);
    for (int64_t j = blockIdx.y; j < n; j += gridDim.y)
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
            const uint64_t z = mix64(sh ^ ((uint64_t)(j0 + j) * 0xD1342543DE82EF95ULL + (uint64_t)(i0 + i)));
            A[j * lda + i] = (double)(z >> 11) * 0x1.0p-53;
        }
}

}  // namespace dhqr
