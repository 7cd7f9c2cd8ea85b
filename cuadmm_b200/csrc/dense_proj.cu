// dense_proj.cu — PSD projection of LARGE blocks (n > 168) with FP64 tensor-core GEMMs only.
//
// Replaces cusolverDnXsyevd + diag scale + cublasDgemm for the reference's "large" blocks
// (src/solver.cu:540-564, 600-644).  A full eigendecomposition is not needed to project: with the
// matrix sign function U = sign(A),   Pi_+(A) = (A + U A) / 2.   U is the limit of an odd polynomial
// iteration  X <- p_k(X) = X (a_k I + b_k X^2 + c_k X^4)  started from X_0 = A / s, s >= ||A||_2:
//   * scale: A_0 = A / ||A||_F; the first product A_0^2 is needed anyway and gives the sharper bound
//     s = ||A_0^2||_F^(1/2) >= ||A_0||_2 for free (its coefficients are folded into step 0);
//   * steps 0..6: degree-5 polynomials that are minimax-optimal for the sign function on [l_k, 1]
//     (each maps [l_k, 1] into [l_{k+1}, 1] with the largest possible l_{k+1}; l_0 = 1e-4 -> 0.944 in seven
//     steps, slope 4.26 at 0 against 1.5 for Newton-Schulz; table from scripts/sign_poly_table.py);
//     p_k <= 1 on [0, 1] and p_k(x) >= x on [0, l_k], so smaller eigenvalues are never pushed back and
//     the conditioning of sign() never gets worse than that of the scaled input — accuracy stays ~1e-14;
//   * then the cubically convergent Newton-Schulz polynomial (15 x - 10 x^3 + 3 x^5) / 8 until
//     ||X^2 - I||_F^2 < 1e-10 (the step after is ~1e-30) or the residual stagnates (exact or tiny zero
//     eigenvalues: what is left unresolved contributes less than its own magnitude, < 1e-10 ||A||).
// A block that is done is frozen (its later products degrade to copies / early exits).  All iterates
// are polynomials in A, hence symmetric: every product is computed on the lower tile triangle only and
// mirrored (n^3 flop each), and every flop is a dense DMMA contraction (mma.sync.m8n8k4.f64 — tcgen05
// has no f64 kind), batched over all large blocks of the plan: 3 products per step, ~30 for a random
// symmetric matrix (the plain Newton-Schulz iteration this replaced needed ~70).  One-stage
// tridiagonalisation would be half BLAS-2 and HBM-bound; this is not.
#include "plan.h"
#include <algorithm>
#include <cmath>

namespace cuadmm {

#ifndef CUADMM_SG_K
#define CUADMM_SG_K 16
#endif
#ifndef CUADMM_SG_STAGES
#define CUADMM_SG_STAGES 4
#endif
// Two tile shapes, one kernel template: WM x WN warps of 32 x 64 each.
//   <2, 1>:  64 x  64 tile,  64 threads, 3 stages ( 52 KB): four CTAs per SM.  Co-resident CTAs overlap one tile's prologue
//            and epilogue with the others' main loops and the finer tiles waste less on the padded edge and the diagonal:
//            measured against the 128-tile shape (scripts/dense_probe.py, ms per projection): 64 x n=200 4.61 -> 4.04,
//            32 x 320 5.66 -> 3.25, 16 x 500 7.33 -> 4.19, 16 x 800 22.2 -> 18.2, 4 x 1200 14.8 -> 14.6, n = 2000 and
//            n = 4000 equal.  Used up to kSmallTileMax;
//   <4, 2>: 128 x 128 tile, 256 threads, 4 stages (135 KB): one CTA per SM, half the operand traffic per flop — kept for
//            the blocks whose operands no longer fit the L2 (n > 2048).
static constexpr int SG_K = CUADMM_SG_K, SG_PAD = 4, SG_STAGES = CUADMM_SG_STAGES, SG_STAGES_SMALL = 3;
template <int WM, int WN> struct SgShape {
    static constexpr int TM = 32 * WM, TN = 64 * WN, THREADS = 32 * WM * WN;
    static constexpr int STAGES = (WM * WN >= 8) ? SG_STAGES : SG_STAGES_SMALL;
    static constexpr size_t SMEM = sizeof(double) * STAGES * SG_K * ((TM + SG_PAD) + (TN + SG_PAD));
    static_assert(TM == TN, "symmetric products need square tiles");
};
using SgBig = SgShape<4, 2>;
using SgSmall = SgShape<2, 1>;
static constexpr int kSmallTileMax = 2048;

struct DenseDesc {
    int64_t off;        // element offset of the block in every n x n pool
    int64_t svec_off;
    int32_t n;
    int32_t pad;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(sa), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, int src_bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(sa), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Grouped symmetric product:  C = alpha * A * B + dshift * I + gamma * D   for symmetric A, B (and
// commuting, so that C is symmetric).  Tile (ti >= tj) of problem blockIdx.y; the strict lower tiles
// are mirrored into the upper triangle.  128 x 128 x 16 tiles, 8 warps (4 x 2), warp tile 32 x 64 =
// 4 x 8 DMMA tiles, double-buffered shared memory, row stride == 4 (mod 16) for conflict-free
// fragment loads.  B is read through its transpose (B symmetric) so both operand loads are coalesced.
// Per-block step-0 scale: alpha *= sc^pa, gamma *= sc^pg with sc = ||A_0^2||_F^(-1/2) read from
// (res base)[0] (pa = pg = 0 otherwise).
struct SignStep {
    int mode;        // 0 plain; 1 X2 = X X, residual ||X2 - I||_F^2 -> res[k]; 3 same, but ||X2||_F^2 -> res[k] (step 0);
                     // 2 middle product (frozen: skip); 4 last product of a step (frozen: copy A -> C)
    int k;           // step index (residual slot)
    int pa, pg;      // powers of the step-0 scale applied to alpha / gamma
    int stag_from;   // residual stagnation may freeze a block once k - 1 >= stag_from
};

__device__ __forceinline__ bool sign_frozen(const double* __restrict__ res, int k, double tolsq, int stag_from) {
    // res[j] = ||X_j^2 - I||_F^2 for j >= 1 (res[0] holds the step-0 scale); a frozen block carries its last
    // residual forward, so it stays frozen
    if (k < 2) return false;
    const double r1 = res[k - 1];
    if (r1 < tolsq) return true;
    if (k - 1 >= stag_from && k >= 3) {
        const double r0 = res[k - 2];
        if (fabs(r0 - r1) <= 1e-13 * r1) return true;
    }
    return false;
}

// work[i] = (block, linear lower-triangular tile index): one CTA per entry
template <int WM, int WN>
__global__ void __launch_bounds__(SgShape<WM, WN>::THREADS) sym_gemm_kernel(const DenseDesc* __restrict__ desc,
        const int2* __restrict__ work,
        const double* __restrict__ Ap, const double* __restrict__ Bp, double* Cp, const double* __restrict__ Dp,
        double alpha, double dshift, double gamma, const int* __restrict__ done_flag,
        SignStep st, double* res_all, int res_stride, double tolsq) {
    using Sh = SgShape<WM, WN>;
    constexpr int SG_M = Sh::TM, SG_N = Sh::TN, SG_THREADS = Sh::THREADS, STAGES = Sh::STAGES;
    if (done_flag && *done_flag) return;
    const int mode = st.mode;
    const int2 wk = work[blockIdx.x];
    const int blk_id = wk.x;
    double* res = res_all ? res_all + (size_t)blk_id * res_stride : nullptr;
    bool frozen = false;
    if (mode != 0) frozen = sign_frozen(res, st.k, tolsq, st.stag_from);
    if ((mode == 1 || mode == 3) && frozen) {
        if (wk.y == 0 && threadIdx.x == 0) res[st.k] = res[st.k - 1];
        return;
    }
    if (mode == 2 && frozen) return;
    if (st.pa | st.pg) {
        const double f2 = res[0];                    // ||A_0^2||_F^2
        const double sc = f2 > 0.0 ? rsqrt(sqrt(f2)) : 0.0;
        double pw = 1.0;
        for (int i = 0; i < st.pa; ++i) pw *= sc;
        alpha *= pw;
        pw = 1.0;
        for (int i = 0; i < st.pg; ++i) pw *= sc;
        gamma *= pw;
    }
    extern __shared__ double sg_smem[];
    double (*As)[SG_K][SG_M + SG_PAD] = reinterpret_cast<double (*)[SG_K][SG_M + SG_PAD]>(sg_smem);
    double (*Bs)[SG_K][SG_N + SG_PAD] = reinterpret_cast<double (*)[SG_K][SG_N + SG_PAD]>(sg_smem + STAGES * SG_K * (SG_M + SG_PAD));
    const DenseDesc d = desc[blk_id];
    const int n = d.n;
    // linear lower-triangular tile index -> (ti, tj), ti >= tj
    const int p = wk.y;
    int ti = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
    while ((ti + 1) * (ti + 2) / 2 <= p) ++ti;
    while (ti * (ti + 1) / 2 > p) --ti;
    const int tj = p - ti * (ti + 1) / 2;
    const double* __restrict__ A = Ap + d.off;
    const double* __restrict__ B = Bp + d.off;
    double* C = Cp + d.off;
    const int64_t ld = n;
    const int m0 = ti * SG_M, n0 = tj * SG_N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp % WM) * 32, wn = (warp / WM) * 64;

    if (mode == 4 && frozen) {
        // converged block: X_{k+1} = X_k (tile copy, mirrored like the product)
        for (int e = tid; e < SG_M * SG_N; e += SG_THREADS) {
            const int gm = m0 + e % SG_M, gn = n0 + e / SG_M;
            if (gm < n && gn < n) {
                const double v = A[gm + gn * ld];
                C[gm + gn * ld] = v;
                if (ti != tj) C[gn + gm * ld] = v;
            }
        }
        return;
    }
    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    // Operand staging: SG_STAGES-deep cp.async ring (global -> shared without a register round trip, zero
    // fill outside the matrix through the src-size operand), one __syncthreads per k-tile.
    // 16-byte copies when every column of the block is 16-byte aligned, 8-byte copies otherwise.
    const bool v16 = ((n & 1) == 0) && ((d.off & 1) == 0);
    auto load_tiles = [&](int buf, int k0) {
        if (v16) {
            // A tile 128 (m) x 16 (k): element (mm, kk) at A[m0+mm + (k0+kk) ld] (contiguous along m); same for B
            // through its transpose: B(kk, nn) = B(nn, kk) at B[n0+nn + (k0+kk) ld]
#pragma unroll
            for (int t = 0; t < (SG_K * SG_M / 2) / SG_THREADS; ++t) {
                const int e = tid + t * SG_THREADS;            // SG_M / 2 16-byte chunks per k-row and operand
                const int mm = (e % (SG_M / 2)) * 2, kk = e / (SG_M / 2);
                const int gk = k0 + kk;
                const int gm = m0 + mm, gn = n0 + mm;
                const bool oka = gm < n && gk < n, okb = gn < n && gk < n;
                cp_async16(&As[buf][kk][mm], oka ? A + gm + gk * ld : A, oka ? 16 : 0);
                cp_async16(&Bs[buf][kk][mm], okb ? B + gn + gk * ld : B, okb ? 16 : 0);
            }
        } else {
#pragma unroll
            for (int t = 0; t < (SG_K * SG_M) / SG_THREADS; ++t) {
                const int e = tid + t * SG_THREADS;
                const int mm = e % SG_M, kk = e / SG_M;
                const int gk = k0 + kk;
                const int gm = m0 + mm, gn = n0 + mm;
                const bool oka = gm < n && gk < n, okb = gn < n && gk < n;
                cp_async8(&As[buf][kk][mm], oka ? A + gm + gk * ld : A, oka ? 8 : 0);
                cp_async8(&Bs[buf][kk][mm], okb ? B + gn + gk * ld : B, okb ? 8 : 0);
            }
        }
    };

    const int nk = (n + SG_K - 1) / SG_K;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_tiles(s, s * SG_K);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt % STAGES;
        cp_async_wait<STAGES - 2>();           // tile kt has landed (this thread's copies) ...
        __syncthreads();                       // ... and everybody's; everybody is also done with tile kt-1's buffer
        if (kt + STAGES - 1 < nk) load_tiles((kt + STAGES - 1) % STAGES, (kt + STAGES - 1) * SG_K);
        cp_async_commit();
#pragma unroll
        for (int kk = 0; kk < SG_K; kk += 4) {
            double a[4], b[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[buf][kk + (lane & 3)][wm + i * 8 + (lane >> 2)];
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = Bs[buf][kk + (lane & 3)][wn + j * 8 + (lane >> 2)];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
    const double* __restrict__ D = Dp ? Dp + d.off : nullptr;
    double rsum = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int gm = m0 + wm + i * 8 + (lane >> 2);
                const int gn = n0 + wn + j * 8 + (lane & 3) * 2 + h;
                // only the lower triangle (gm >= gn) is kept and mirrored, also inside diagonal tiles:
                // the iterates must stay EXACTLY symmetric (B is read through its transpose), an
                // antisymmetric rounding residue of 1e-16 otherwise grows to 1e-9 over the iteration
                if (gm < n && gn < n && gm >= gn) {
                    if (mode == 1 || mode == 3) {
                        const double t = acc[i][j][h] - ((gm == gn && mode == 1) ? 1.0 : 0.0);
                        rsum = fma(gm != gn ? 2.0 * t : t, t, rsum);
                    }
                    double v = alpha * acc[i][j][h];
                    if (gm == gn) v += dshift;
                    if (D) v += gamma * D[gm + gn * ld];
                    C[gm + gn * ld] = v;
                    if (gm != gn) C[gn + gm * ld] = v;
                }
            }
    if (mode == 1 || mode == 3) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
        if (lane == 0) atomicAdd(res + st.k, rsum);
    }
}

// svec -> dense symmetric A (both triangles), per-block sum of squares (= ||A||_F^2, svec is an isometry)
__global__ void dense_load_kernel(const DenseDesc* __restrict__ desc, const double* __restrict__ svec, double* A,
                                  double* fro2, const int* __restrict__ done_flag) {
    if (done_flag && *done_flag) return;
    const DenseDesc d = desc[blockIdx.y];
    const int n = d.n;
    const double* x = svec + d.svec_off;
    double* M = A + d.off;
    double f = 0.0;
    for (int c = blockIdx.x; c < n; c += gridDim.x) {
        const int64_t base = (int64_t)c * (c + 1) / 2;
        for (int r = threadIdx.x; r <= c; r += blockDim.x) {
            const double s = x[base + r];
            f = fma(s, s, f);
            const double v = (r == c) ? s : s * CUADMM_SQRT2INV;
            M[(int64_t)n * c + r] = v;
            M[(int64_t)n * r + c] = v;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) f += __shfl_xor_sync(0xffffffffu, f, o);
    if ((threadIdx.x & 31) == 0 && f != 0.0) atomicAdd(fro2 + blockIdx.y, f);
}

// X = A / ||A||_F  (0 for a zero block)
__global__ void dense_scale_kernel(const DenseDesc* __restrict__ desc, const double* __restrict__ A, double* X,
                                   const double* __restrict__ fro2, const int* __restrict__ done_flag) {
    if (done_flag && *done_flag) return;
    const DenseDesc d = desc[blockIdx.y];
    const int64_t nn = (int64_t)d.n * d.n;
    const double f = fro2[blockIdx.y];
    const double inv = (f > 0.0 && f < 1.7e308) ? rsqrt(f) : (f == 0.0 ? 0.0 : f - f);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nn; e += (int64_t)gridDim.x * blockDim.x)
        X[d.off + e] = A[d.off + e] * inv;
}

// dense P -> svec (lower/upper averaged), optional fused S / SmC epilogue
__global__ void dense_store_kernel(const DenseDesc* __restrict__ desc, const double* __restrict__ P, double* out,
                                   ProjEpilogue epi, const int* __restrict__ done_flag) {
    if (done_flag && *done_flag) return;
    const DenseDesc d = desc[blockIdx.y];
    const int n = d.n;
    const double* M = P + d.off;
    double sig = 1.0;
    if (epi.X) sig = *epi.sig_ptr;
    for (int c = blockIdx.x; c < n; c += gridDim.x) {
        const int64_t base = (int64_t)c * (c + 1) / 2;
        for (int r = threadIdx.x; r <= c; r += blockDim.x) {
            const double v = 0.5 * (M[(int64_t)n * c + r] + M[(int64_t)n * r + c]);
            const double o = (r == c) ? v : v * CUADMM_SQRT2;
            const int64_t gi = d.svec_off + base + r;
            out[gi] = o;
            if (epi.X) {
                const double Sv = (o - epi.X[gi]) / sig - epi.Rd1[gi];
                epi.S[gi] = Sv;
                epi.SmC[gi] = Sv - epi.Cd[gi];
            }
        }
    }
}

static constexpr int kMaxNsSteps = 40;
static constexpr double kNsTolSq = 1e-10;   // freeze once ||X^2 - I||_F^2 < 1e-10 (cubic: the step after is ~1e-30)
// minimax degree-5 sign polynomials a x + b x^3 + c x^5 on [l_k, 1]  (scripts/sign_poly_table.py, l_0 = 1e-4)
static constexpr int kSignTable = 7;
static const double kSignPoly[kSignTable][3] = {
    {4.2567538552083883, -12.637529108386531, 9.3807751804785546},   // [1.000e-04, 1] -> [0.000426, 1]
    {4.2554599297572135, -12.626635057799284, 9.3711750746066897},   // [4.257e-04, 1] -> [0.001811, 1]
    {4.249950912774862, -12.580323192655207, 9.3303722673722636},    // [1.811e-03, 1] -> [0.007698, 1]
    {4.226491089776534, -12.384384556067777, 9.1578934252368036},    // [7.698e-03, 1] -> [0.032531, 1]
    {4.1267873247419073, -11.574501430286471, 8.4477140393475434},   // [3.253e-02, 1] -> [0.133848, 1]
    {3.7225306718083893, -8.6538466523498734, 5.9313159658509376},   // [1.338e-01, 1] -> [0.477758, 1]
    {2.658148855151357, -3.3824715696999976, 1.7243227088946445},    // [4.778e-01, 1] -> [0.944015, 1]
};
static const double kNs5[3] = {15.0 / 8.0, -10.0 / 8.0, 3.0 / 8.0};

struct DensePart {
    int device = -1;
    std::vector<DenseDesc> h_desc;
    DevBuf<DenseDesc> d_desc;
    DevBuf<double> A, X, Y, Z;      // four n x n pools
    DevBuf<double> fro2;
    DevBuf<double> res;             // (kMaxNsSteps + 1) residual slots per block, block-major
    int nmax = 0;
    DevBuf<int2> work_big, work_small;   // (block, tile) lists of the two tile shapes
    int n_big = 0, n_small = 0;
};

// lower-triangular tile lists of a set of blocks: blocks up to small_max use the 64-tile shape
static void sym_work_lists(const std::vector<DenseDesc>& desc, int small_max, std::vector<int2>& big, std::vector<int2>& small) {
    for (int b = 0; b < (int)desc.size(); ++b) {
        const bool sm = desc[b].n <= small_max;
        const int tile = sm ? SgSmall::TM : SgBig::TM;
        const int T = (desc[b].n + tile - 1) / tile;
        for (int p = 0; p < T * (T + 1) / 2; ++p) (sm ? small : big).push_back(make_int2(b, p));
    }
}

static int small_tile_max() {
    int v = kSmallTileMax;
    if (const char* e = getenv("CUADMM_SG_SMALL_MAX")) v = atoi(e);     // tuning: 0 = every block on 128-tiles
    return v;
}

static void sym_gemm_attrs() {
    CUADMM_CUDA(cudaFuncSetAttribute((const void*)sym_gemm_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SgBig::SMEM));
    CUADMM_CUDA(cudaFuncSetAttribute((const void*)sym_gemm_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SgSmall::SMEM));
}

// one grouped product over both tile lists; returns the number of launches
static int sym_gemm_launch(const DenseDesc* desc, const int2* wbig, int nbig, const int2* wsmall, int nsmall, const double* A,
                           const double* B, double* C, const double* Dm, double alpha, double dshift, double gamma,
                           const int* done_flag, const SignStep& g, double* res, int rs, double tolsq, cudaStream_t st) {
    int launches = 0;
    if (nbig) {
        sym_gemm_kernel<4, 2><<<nbig, SgBig::THREADS, SgBig::SMEM, st>>>(desc, wbig, A, B, C, Dm, alpha, dshift, gamma, done_flag, g, res, rs, tolsq);
        ++launches;
    }
    if (nsmall) {
        sym_gemm_kernel<2, 1><<<nsmall, SgSmall::THREADS, SgSmall::SMEM, st>>>(desc, wsmall, A, B, C, Dm, alpha, dshift, gamma, done_flag, g, res, rs, tolsq);
        ++launches;
    }
    return launches;
}

}  // namespace cuadmm

using namespace cuadmm;

void dense_part_destroy(cuadmm::DensePart* p) { delete p; }

cuadmm::DensePart* dense_part_create(int device, const std::vector<int32_t>& blk, const std::vector<int64_t>& svec_off,
                                     const std::vector<int64_t>& which) {
    std::unique_ptr<DensePart> D(new DensePart());
    D->device = device;
    int64_t off = 0;
    for (int64_t k : which) {
        DenseDesc d;
        d.off = off; d.svec_off = svec_off[k]; d.n = blk[k]; d.pad = 0;
        off += (int64_t)blk[k] * blk[k];
        D->nmax = std::max(D->nmax, (int)blk[k]);
        D->h_desc.push_back(d);
    }
    // largest first: their tile grids dominate
    std::stable_sort(D->h_desc.begin(), D->h_desc.end(), [](const DenseDesc& a, const DenseDesc& b) { return a.n > b.n; });
    if (device < 0 || which.empty()) return D.release();
    D->d_desc.upload(D->h_desc);
    D->A.alloc(off); D->X.alloc(off); D->Y.alloc(off); D->Z.alloc(off);
    D->fro2.alloc((int64_t)which.size());
    D->res.alloc((int64_t)which.size() * (kMaxNsSteps + 1));
    std::vector<int2> big, small;
    sym_work_lists(D->h_desc, small_tile_max(), big, small);
    D->n_big = (int)big.size(); D->n_small = (int)small.size();
    if (!big.empty()) D->work_big.upload(big);
    if (!small.empty()) D->work_small.upload(small);
    sym_gemm_attrs();
    return D.release();
}

// returns the number of launches
int dense_part_project(cuadmm::DensePart* D, const double* Xb, double* Xproj, cudaStream_t st,
                       const cuadmm::ProjEpilogue* epi, const int* done_flag) {
    const int nb = (int)D->h_desc.size();
    if (nb == 0) return 0;
    int launches = 0;
    CUADMM_CUDA(cudaMemsetAsync(D->fro2.p, 0, sizeof(double) * nb, st));
    dim3 gl(std::min(std::max(D->nmax / 8, 1), 512), nb);
    dense_load_kernel<<<gl, 256, 0, st>>>(D->d_desc.p, Xb, D->A.p, D->fro2.p, done_flag);
    dim3 gs(std::min((int)(((int64_t)D->nmax * D->nmax + 255) / 256), 2048), nb);
    dense_scale_kernel<<<gs, 256, 0, st>>>(D->d_desc.p, D->A.p, D->X.p, D->fro2.p, done_flag);
    launches += 2;
    double* X = D->X.p; double* Y = D->Y.p; double* Z = D->Z.p;
    const int rs = kMaxNsSteps + 1;
    CUADMM_CUDA(cudaMemsetAsync(D->res.p, 0, sizeof(double) * (size_t)nb * rs, st));
    for (int k = 0; k < kMaxNsSteps; ++k) {
        const double* co = k < kSignTable ? kSignPoly[k] : kNs5;
        const int s0 = (k == 0);
        // Y = X X (+ residual of X_k; step 0: ||X_0^2||_F^2, the scale) ; Z = c Y Y + b Y ; Y = X Z + a X ; swap(X, Y)
        SignStep g1{s0 ? 3 : 1, k, 0, 0, kSignTable + 1}, g2{2, k, s0 ? 4 : 0, s0 ? 2 : 0, kSignTable + 1},
                 g3{4, k, s0 ? 1 : 0, s0 ? 1 : 0, kSignTable + 1};
        auto prod = [&](const double* A_, const double* B_, double* C_, const double* D_, double al, double ga, const SignStep& g) {
            launches += sym_gemm_launch(D->d_desc.p, D->work_big.p, D->n_big, D->work_small.p, D->n_small, A_, B_, C_, D_, al, 0.0, ga,
                                        done_flag, g, D->res.p, rs, kNsTolSq, st);
        };
        prod(X, X, Y, nullptr, 1.0, 0.0, g1);
        prod(Y, Y, Z, Y, co[2], co[1], g2);
        prod(X, Z, Y, X, 1.0, co[0], g3);
        std::swap(X, Y);
    }
    // P = (U A + A) / 2 with U = sign(A) in X
    SignStep plain{0, 0, 0, 0, 0};
    launches += sym_gemm_launch(D->d_desc.p, D->work_big.p, D->n_big, D->work_small.p, D->n_small, X, D->A.p, Y, D->A.p, 0.5, 0.0, 0.5,
                                done_flag, plain, nullptr, rs, 0.0, st);
    double* Pout = Y;
    ProjEpilogue e;
    if (epi) e = *epi; else { e.X = nullptr; e.Rd1 = nullptr; e.Cd = nullptr; e.S = nullptr; e.SmC = nullptr; e.sig_ptr = nullptr; }
    dense_store_kernel<<<gl, 256, 0, st>>>(D->d_desc.p, Pout, Xproj, e, done_flag);
    launches += 1;
    CUADMM_CUDA(cudaGetLastError());
    return launches;
}

// test hook (not in the public header): C = alpha * A * B + dshift * I for one n x n symmetric pair
extern "C" int cuadmm_debug_sym_gemm(int n, const double* hA, const double* hB, double* hC, double alpha, double dshift) {
    return cuadmm::guarded([&] {
        const int64_t nn = (int64_t)n * n;
        DevBuf<double> A(nn), B(nn), C(nn);
        A.upload(hA, nn); B.upload(hB, nn);
        DenseDesc d; d.off = 0; d.svec_off = 0; d.n = n; d.pad = 0;
        DevBuf<DenseDesc> dd(1);
        CUADMM_CUDA(cudaMemcpy(dd.p, &d, sizeof d, cudaMemcpyHostToDevice));
        sym_gemm_attrs();
        std::vector<int2> big, small;
        sym_work_lists(std::vector<DenseDesc>(1, d), small_tile_max(), big, small);
        DevBuf<int2> wb, ws;
        if (!big.empty()) wb.upload(big);
        if (!small.empty()) ws.upload(small);
        SignStep plain{0, 0, 0, 0, 0};
        sym_gemm_launch(dd.p, wb.p, (int)big.size(), ws.p, (int)small.size(), A.p, B.p, C.p, nullptr, alpha, dshift, 0.0, nullptr,
                        plain, nullptr, 1, 0.0, 0);
        CUADMM_CUDA(cudaGetLastError());
        C.download(hC, nn);
        CUADMM_CUDA(cudaDeviceSynchronize());
    });
}
