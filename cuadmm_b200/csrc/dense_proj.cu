// dense_proj.cu — PSD projection of LARGE blocks (n > 168) with FP64 tensor-core GEMMs only.
//
// Replaces cusolverDnXsyevd + diag scale + cublasDgemm for the reference's "large" blocks
// (src/solver.cu:540-564, 600-644).  A full eigendecomposition is not needed to project: with the
// matrix sign function U = sign(A),   Pi_+(A) = (A + U A) / 2.   U is computed by the Newton-Schulz
// iteration  X <- X (3 I - X^2) / 2,  X_0 = A / ||A||_F  (spectrum in [-1, 1]; monotone on it, so no
// eigenvalue is ever pushed towards 0 and the conditioning of sign() is that of A itself — a scaled
// variant that maps the top of the spectrum onto a worst-case lower bound was measured to lose 5+
// digits).  Convergence is quadratic; every step accumulates ||X^2 - I||_F^2 in its first product and
// a block whose residual fell below 1e-14 is frozen (its later products degrade to copies), at most
// 60 steps.  Eigenvalues with |lambda| < ~1e-10 ||A||_F may stay unconverged; they contribute less
// than |lambda| to the projection.  All iterates are polynomials in A, hence symmetric: every product
// is computed on the lower tile triangle only and mirrored (n^3 flop each), and every flop is a dense
// DMMA contraction (mma.sync.m8n8k4.f64 — tcgen05 has no f64 kind), batched over all large blocks of
// the plan.  One-stage tridiagonalisation would be half BLAS-2 and HBM-bound; this is not.
#include "plan.h"
#include <algorithm>
#include <cmath>

namespace cuadmm {

static constexpr int SG_M = 128, SG_N = 128, SG_K = 16, SG_PAD = 4, SG_THREADS = 256;

struct DenseDesc {
    int64_t off;        // element offset of the block in every n x n pool
    int64_t svec_off;
    int32_t n;
    int32_t pad;
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Grouped symmetric product:  C = alpha * A * B + dshift * I + gamma * D   for symmetric A, B (and
// commuting, so that C is symmetric).  Tile (ti >= tj) of problem blockIdx.y; the strict lower tiles
// are mirrored into the upper triangle.  128 x 128 x 16 tiles, 8 warps (4 x 2), warp tile 32 x 64 =
// 4 x 8 DMMA tiles, double-buffered shared memory, row stride == 4 (mod 16) for conflict-free
// fragment loads.  B is read through its transpose (B symmetric) so both operand loads are coalesced.
__global__ void __launch_bounds__(SG_THREADS) sym_gemm_kernel(const DenseDesc* __restrict__ desc,
        const double* __restrict__ Ap, const double* __restrict__ Bp, double* Cp, const double* __restrict__ Dp,
        double alpha, double dshift, double gamma, const int* __restrict__ done_flag,
        int mode, const double* __restrict__ res_prev, double* res_cur, int res_stride, double tolsq) {
    // mode 1: first product of a step  (C = 1.5 I - 0.5 X X; accumulates ||X X - I||_F^2 into res_cur)
    // mode 2: second product of a step (C = X Y), a frozen block copies X instead
    // mode 0: plain
    if (done_flag && *done_flag) return;
    bool frozen = false;
    if (mode != 0 && res_prev) frozen = res_prev[blockIdx.y * res_stride] < tolsq;
    if (mode == 1 && frozen) {
        if (blockIdx.x == 0 && threadIdx.x == 0) res_cur[blockIdx.y * res_stride] = res_prev[blockIdx.y * res_stride];
        return;
    }
    extern __shared__ double sg_smem[];
    double (*As)[SG_K][SG_M + SG_PAD] = reinterpret_cast<double (*)[SG_K][SG_M + SG_PAD]>(sg_smem);
    double (*Bs)[SG_K][SG_N + SG_PAD] = reinterpret_cast<double (*)[SG_K][SG_N + SG_PAD]>(sg_smem + 2 * SG_K * (SG_M + SG_PAD));
    const DenseDesc d = desc[blockIdx.y];
    const int n = d.n;
    const int T = (n + SG_M - 1) / SG_M;
    // linear lower-triangular tile index -> (ti, tj), ti >= tj
    const int p = blockIdx.x;
    if (p >= T * (T + 1) / 2) return;
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
    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 64;

    if (mode == 2 && frozen) {
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

    auto load_tiles = [&](int buf, int k0) {
        // A tile 128 (m) x 16 (k): element (mm, kk) at A[m0+mm + (k0+kk) ld]   (coalesced along m)
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int e = tid + t * SG_THREADS;
            const int mm = e % SG_M, kk = e / SG_M;
            const int gm = m0 + mm, gk = k0 + kk;
            As[buf][kk][mm] = (gm < n && gk < n) ? A[gm + gk * ld] : 0.0;
        }
        // B tile 16 (k) x 128 (n): B(kk, nn) = B(nn, kk) (symmetric) at B[n0+nn + (k0+kk) ld]
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int e = tid + t * SG_THREADS;
            const int nn = e % SG_N, kk = e / SG_N;
            const int gn = n0 + nn, gk = k0 + kk;
            Bs[buf][kk][nn] = (gn < n && gk < n) ? B[gn + gk * ld] : 0.0;
        }
    };

    const int nk = (n + SG_K - 1) / SG_K;
    load_tiles(0, 0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles(buf ^ 1, (kt + 1) * SG_K);
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
        __syncthreads();
    }
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
                    if (mode == 1) {
                        const double t = acc[i][j][h] - (gm == gn ? 1.0 : 0.0);
                        rsum = fma(gm != gn ? 2.0 * t : t, t, rsum);
                    }
                    double v = alpha * acc[i][j][h];
                    if (gm == gn) v += dshift;
                    if (D) v += gamma * D[gm + gn * ld];
                    C[gm + gn * ld] = v;
                    if (gm != gn) C[gn + gm * ld] = v;
                }
            }
    if (mode == 1) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
        if (lane == 0) atomicAdd(res_cur + blockIdx.y * res_stride, rsum);
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

static constexpr int kMaxNsSteps = 60;
static constexpr double kNsTolSq = 1e-14;   // freeze once ||X^2 - I||_F^2 < 1e-14 (the step after is ~1e-28)

struct DensePart {
    int device = -1;
    std::vector<DenseDesc> h_desc;
    DevBuf<DenseDesc> d_desc;
    DevBuf<double> A, X, Y, Z;      // four n x n pools
    DevBuf<double> fro2;
    DevBuf<double> res;             // (kMaxNsSteps + 1) residual slots per block, block-major
    int nmax = 0;
    size_t smem = 0;
};

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
    D->smem = sizeof(double) * 2 * SG_K * ((SG_M + SG_PAD) + (SG_N + SG_PAD));
    CUADMM_CUDA(cudaFuncSetAttribute((const void*)sym_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D->smem));
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
    const int T = (D->nmax + SG_M - 1) / SG_M;
    dim3 gg(T * (T + 1) / 2, nb);
    double* X = D->X.p; double* Xn = D->Z.p;
    const int rs = kMaxNsSteps + 1;
    // res[.., 0] = +inf-like (never frozen before the first step)
    CUADMM_CUDA(cudaMemsetAsync(D->res.p, 0, sizeof(double) * (size_t)nb * rs, st));
    for (int k = 0; k < kMaxNsSteps; ++k) {
        // Y = 1.5 I - 0.5 X X (+ residual of step k) ;  X' = X Y   (frozen blocks: copy)
        const double* rprev = (k == 0) ? nullptr : D->res.p + (k - 1);
        sym_gemm_kernel<<<gg, SG_THREADS, D->smem, st>>>(D->d_desc.p, X, X, D->Y.p, nullptr, -0.5, 1.5, 0.0, done_flag,
                                                          1, rprev, D->res.p + k, rs, kNsTolSq);
        sym_gemm_kernel<<<gg, SG_THREADS, D->smem, st>>>(D->d_desc.p, X, D->Y.p, Xn, nullptr, 1.0, 0.0, 0.0, done_flag,
                                                          2, rprev, nullptr, rs, kNsTolSq);
        std::swap(X, Xn);
        launches += 2;
    }
    // P = (U A + A) / 2 with U = sign(A) in X
    sym_gemm_kernel<<<gg, SG_THREADS, D->smem, st>>>(D->d_desc.p, X, D->A.p, D->Y.p, D->A.p, 0.5, 0.0, 0.5, done_flag,
                                                      0, nullptr, nullptr, rs, 0.0);
    ProjEpilogue e;
    if (epi) e = *epi; else { e.X = nullptr; e.Rd1 = nullptr; e.Cd = nullptr; e.S = nullptr; e.SmC = nullptr; e.sig_ptr = nullptr; }
    dense_store_kernel<<<gl, 256, 0, st>>>(D->d_desc.p, D->Y.p, Xproj, e, done_flag);
    launches += 2;
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
        const size_t smem = sizeof(double) * 2 * SG_K * ((SG_M + SG_PAD) + (SG_N + SG_PAD));
        CUADMM_CUDA(cudaFuncSetAttribute((const void*)sym_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int T = (n + SG_M - 1) / SG_M;
        sym_gemm_kernel<<<dim3(T * (T + 1) / 2, 1), SG_THREADS, smem>>>(dd.p, A.p, B.p, C.p, nullptr, alpha, dshift, 0.0, nullptr,
                                                                        0, nullptr, nullptr, 1, 0.0);
        CUADMM_CUDA(cudaGetLastError());
        C.download(hC, nn);
        CUADMM_CUDA(cudaDeviceSynchronize());
    });
}
