// ysolve.cu — y = (A A^T + eps I)^-1 rhs on the device.
//
// Replaces the reference's per-iteration host path (src/solver.cu:487-500, 704-717):
//   perform_permutation -> cudaDeviceSynchronize -> D2H -> cholmod_solve2 (host, simplicial LDL^T)
//   -> H2D -> perform_permutation
// with: gather-permute fused into a synchronisation-free sparse forward sweep, a dense GEMV pair
// on the explicitly inverted trailing block of the factor (where the elimination DAG degenerates
// into a chain), a sparse backward sweep with the scatter-permute fused in.  Nothing leaves the
// GPU and there is no host synchronisation.
#include "ysolve.h"
#include "dense.h"
#include <algorithm>
#include <numeric>
#include <stdlib.h>

namespace cuadmm {

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

static constexpr int kTriThreads = 256;
static constexpr int kLongRowNnz = 24;     // rows with more dependencies get a whole warp

// One launch = one whole triangular sweep, level-synchronous without leaving the kernel.
// Work is cut into warp-slots: a slot is either 8 short rows (4 lanes each) or 1 long row (32 lanes),
// all of one dependency level.  A warp may start a slot of level l once counters[l-1] has reached the
// number of slots of level l-1 (acquire poll by lane 0), computes its rows with ordinary coalesced
// loads (no per-entry flag traffic), publishes x with a fence and bumps counters[l].  All CTAs are
// co-resident (grid from the occupancy API) and every warp visits its slots in level order, so the
// sweep cannot deadlock.
__global__ void __launch_bounds__(kTriThreads) tri_level_kernel(
        int64_t n_slots, const int32_t* __restrict__ slot_rows, const int32_t* __restrict__ slot_info,
        const int32_t* __restrict__ level_slots, const int64_t* __restrict__ ptr,
        const int32_t* __restrict__ dep, const double* __restrict__ val, const double* __restrict__ inv_diag,
        const double* __restrict__ rhs, const int32_t* __restrict__ rhs_gather,
        double* x, int* counters, double* out_scatter, const int32_t* __restrict__ out_perm,
        const int* __restrict__ done_flag, unsigned backoff_ns) {
    if (done_flag && *done_flag) return;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * kTriThreads + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * kTriThreads) >> 5;
    for (int64_t s = warp; s < n_slots; s += nwarps) {
        const int32_t info = slot_info[s];
        const int lvl = info & 0x3fffffff;
        const int is_long = info >> 30;
        if (lvl > 0) {
            if (lane == 0) {
                const int target = level_slots[lvl - 1];
                while (ld_acquire_gpu(counters + lvl - 1) < target) { if (backoff_ns) __nanosleep(backoff_ns); }
            }
            __syncwarp();
        }
        double acc = 0.0;
        int32_t u;
        if (is_long) {
            u = slot_rows[8 * s];
            const int64_t p1 = ptr[u + 1];
#pragma unroll 4
            for (int64_t p = ptr[u] + lane; p < p1; p += 32) acc = fma(val[p], __ldcg(x + dep[p]), acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        } else {
            u = slot_rows[8 * s + (lane >> 2)];
            if (u >= 0) {
                const int64_t p1 = ptr[u + 1];
                for (int64_t p = ptr[u] + (lane & 3); p < p1; p += 4) acc = fma(val[p], __ldcg(x + dep[p]), acc);
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        }
        const bool writer = is_long ? (lane == 0) : ((lane & 3) == 0 && u >= 0);
        if (writer) {
            const double r = rhs_gather ? rhs[rhs_gather[u]] : rhs[u];
            const double v = (r - acc) * inv_diag[u];
            __stcg(x + u, v);
            if (out_scatter) out_scatter[out_perm[u]] = v;
            __threadfence();
        }
        __syncwarp();
        if (lane == 0) atomicAdd(counters + lvl, 1);
    }
}

// dense tail: out[i] = sum_j T[i, j] * in[j] for a row-major r x r matrix of which only the
// lower (lower=true) or upper triangle is non-zero.  One warp per row, coalesced.
__global__ void __launch_bounds__(256) tail_gemv_kernel(int64_t r, const double* __restrict__ T, const double* __restrict__ in,
                                                        double* out, int lower, double* out_scatter,
                                                        const int32_t* __restrict__ out_perm, int64_t perm_base,
                                                        const int* __restrict__ done_flag) {
    if (done_flag && *done_flag) return;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= r) return;
    const double* Ti = T + row * r;
    const int64_t j0 = lower ? 0 : row, j1 = lower ? row + 1 : r;
    double acc = 0.0;
    for (int64_t j = j0 + lane; j < j1; j += 32) acc = fma(Ti[j], in[j], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        out[row] = acc;
        if (out_scatter) out_scatter[out_perm[perm_base + row]] = acc;
    }
}

static void launch_sweep(const TriSweep& S, const double* rhs, const int32_t* gather, double* x, int* counters,
                         double* out_scatter, const int32_t* out_perm, const int* done, cudaStream_t st) {
    if (S.n_slots == 0) return;
    tri_level_kernel<<<S.grid, kTriThreads, 0, st>>>(S.n_slots, S.slot_rows.p, S.slot_info.p, S.level_slots.p, S.ptr.p,
        S.dep.p, S.val.p, S.inv_diag.p, rhs, gather, x, counters, out_scatter, out_perm, done, S.backoff_ns);
    CUADMM_CUDA(cudaGetLastError());
}

static int sweep_max_grid(int device) {
    int per_sm = 0, sms = 0;
    CUADMM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tri_level_kernel, kTriThreads, 0));
    CUADMM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    return std::max(1, per_sm) * sms;
}

// host: pull structure -> device sweep
struct HostSweep {
    std::vector<int64_t> ptr;
    std::vector<int32_t> dep;
    std::vector<double> val, inv_diag;
    std::vector<int32_t> level;   // per unknown
    int64_t n = 0;                // unknown id space
    std::vector<int32_t> unknowns;  // ids actually solved by the sweep
};

static void upload_sweep(const HostSweep& H, TriSweep& S, int device) {
    S.n_unknowns = (int64_t)H.unknowns.size();
    S.nnz = (int64_t)H.dep.size();
    int maxlev = -1;
    for (int32_t u : H.unknowns) maxlev = std::max(maxlev, (int)H.level[u]);
    S.levels = maxlev + 1;
    // bucket by level, short rows first (8 per warp-slot), then long rows (1 per warp-slot)
    std::vector<std::vector<int32_t>> shorts(S.levels), longs(S.levels);
    for (int32_t u : H.unknowns) {
        const int64_t len = H.ptr[u + 1] - H.ptr[u];
        (len > kLongRowNnz ? longs : shorts)[H.level[u]].push_back(u);
    }
    std::vector<int32_t> slot_rows, slot_info, level_slots(std::max(S.levels, 1), 0);
    for (int l = 0; l < S.levels; ++l) {
        int cnt = 0;
        // heavier long rows first so the level's tail is short
        std::stable_sort(longs[l].begin(), longs[l].end(), [&](int32_t a, int32_t b) {
            return H.ptr[a + 1] - H.ptr[a] > H.ptr[b + 1] - H.ptr[b]; });
        for (int32_t u : longs[l]) {
            slot_rows.push_back(u);
            for (int t = 1; t < 8; ++t) slot_rows.push_back(-1);
            slot_info.push_back(l | (1 << 30));
            ++cnt;
        }
        for (size_t t = 0; t < shorts[l].size(); t += 8) {
            for (size_t q = t; q < t + 8; ++q) slot_rows.push_back(q < shorts[l].size() ? shorts[l][q] : -1);
            slot_info.push_back(l);
            ++cnt;
        }
        level_slots[l] = cnt;
    }
    S.n_slots = (int64_t)slot_info.size();
    S.ptr.upload(H.ptr);
    std::vector<int32_t> dep(H.dep); if (dep.empty()) dep.push_back(0);
    std::vector<double> val(H.val); if (val.empty()) val.push_back(0.0);
    S.dep.upload(dep); S.val.upload(val);
    S.inv_diag.upload(H.inv_diag);
    if (slot_rows.empty()) { slot_rows.assign(8, -1); slot_info.push_back(0); }
    S.slot_rows.upload(slot_rows); S.slot_info.upload(slot_info); S.level_slots.upload(level_slots);
    int maxgrid = sweep_max_grid(device);
    {   // thousands of warps polling one level counter turn it into an L2 hot spot: 2 CTAs per SM measured best
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        maxgrid = std::min(maxgrid, 2 * sms);
    }
    if (const char* e = getenv("CUADMM_SWEEP_MAXGRID")) maxgrid = std::max(1, std::min(maxgrid, atoi(e)));
    if (const char* e = getenv("CUADMM_SWEEP_BACKOFF_NS")) S.backoff_ns = (unsigned)atoi(e);
    const int64_t need = (S.n_slots * 32 + kTriThreads - 1) / kTriThreads;
    S.grid = (int)std::max<int64_t>(1, std::min<int64_t>(need, maxgrid));
}

// choose the rows that go to the dense tail: the deep, narrow end of the elimination DAG.
static int64_t choose_tail(const CholFactor& F, std::vector<int32_t>& level_out) {
    const int64_t n = F.n;
    std::vector<int32_t> lev(n, 0);
    for (int64_t j = 0; j < n; ++j)
        for (int64_t p = F.Lp[j] + 1; p < F.Lp[j + 1]; ++p) lev[F.Li[p]] = std::max(lev[F.Li[p]], lev[j] + 1);
    level_out = lev;
    int depth = 0;
    for (int64_t i = 0; i < n; ++i) depth = std::max(depth, lev[i] + 1);
    int64_t max_tail = 6144;
    int min_depth = 64;
    if (const char* e = getenv("CUADMM_YSOLVE_MAX_TAIL")) max_tail = atoll(e);
    if (const char* e = getenv("CUADMM_YSOLVE_MIN_DEPTH")) min_depth = atoi(e);
    if (depth <= min_depth || max_tail <= 0) return depth + 1;       // no tail
    std::vector<int64_t> cnt(depth + 1, 0);
    for (int64_t i = 0; i < n; ++i) cnt[lev[i]]++;
    // smallest cut >= min_depth/2 with count(level >= cut) <= max_tail
    int64_t above = 0;
    int cut = depth;
    for (int l = depth - 1; l >= min_depth / 2; --l) {
        if (above + cnt[l] > max_tail) break;
        above += cnt[l];
        cut = l;
    }
    if (above < 32) return depth + 1;   // a tiny tail is not worth a dense stage
    return cut;
}

cuadmm_ysolve_s* ysolve_create(int64_t m, int64_t vec_len, int64_t nnz, const int32_t* rowptr, const int32_t* colind,
                               const double* val, double eps, int device) {
    CUADMM_REQUIRE(m >= 0 && vec_len >= 0 && nnz >= 0, "negative dimension");
    CUADMM_REQUIRE(rowptr && (nnz == 0 || (colind && val)), "null argument");
    CUADMM_REQUIRE(rowptr[0] == 0 && rowptr[m] == nnz, "rowptr does not span nnz");
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        throw Error(CUADMM_ENODEVICE, "no CUDA device available; the y-solve has no CPU fallback");
    }
    CUADMM_REQUIRE(device >= 0 && device < cnt, "device index out of range");
    std::unique_ptr<cuadmm_ysolve_s> Y(new cuadmm_ysolve_s());
    Y->device = device; Y->m = m;
    DeviceGuard g(device);

    // ---- host analysis
    SymCsc M = form_aat(m, vec_len, rowptr, colind, val, eps);
    Y->nnz_aat = M.p[m];
    std::vector<int32_t> perm0 = min_degree_order(M);
    CholFactor F;
    chol_symbolic(M, perm0, F, nullptr);
    std::vector<int32_t> lev;
    const int64_t cut = choose_tail(F, lev);
    std::vector<int32_t> perm1; perm1.reserve(m);
    for (int64_t k = 0; k < m; ++k) if (lev[k] < cut) perm1.push_back(F.perm[k]);
    const int64_t n_lead = (int64_t)perm1.size();
    for (int64_t k = 0; k < m; ++k) if (lev[k] >= cut) perm1.push_back(F.perm[k]);
    const int64_t n_tail = m - n_lead;
    SymCsc C;
    if (n_tail > 0) {
        F = CholFactor();
        chol_symbolic(M, perm1, F, &C);
    } else {
        std::vector<int32_t> ip;
        chol_symbolic(M, perm0, F, &C);
    }
    chol_numeric(C, F, n_lead);
    Y->n_lead = n_lead; Y->n_tail = n_tail;
    Y->nnz_L = F.nnz();
    Y->n_deficient = F.n_deficient;
    Y->h_perm = F.perm;
    Y->perm.upload(F.perm);

    // ---- forward sweep: rows of L (lead rows solve, tail rows accumulate t2 = b2 - L21 z1)
    {
        HostSweep H;
        H.n = m;
        H.ptr.assign(m + 1, 0);
        for (int64_t j = 0; j < n_lead; ++j)
            for (int64_t p = F.Lp[j] + 1; p < F.Lp[j + 1]; ++p) H.ptr[F.Li[p] + 1]++;
        for (int64_t i = 0; i < m; ++i) H.ptr[i + 1] += H.ptr[i];
        H.dep.resize(H.ptr[m]); H.val.resize(H.ptr[m]);
        std::vector<int64_t> nx(H.ptr.begin(), H.ptr.end() - 1);
        for (int64_t j = 0; j < n_lead; ++j)
            for (int64_t p = F.Lp[j] + 1; p < F.Lp[j + 1]; ++p) {
                const int64_t q = nx[F.Li[p]]++;
                H.dep[q] = (int32_t)j; H.val[q] = F.Lx[p];
            }
        H.inv_diag.assign(m, 1.0);
        for (int64_t i = 0; i < n_lead; ++i) H.inv_diag[i] = 1.0 / F.Lx[F.Lp[i]];
        H.level.assign(m, 0);
        for (int64_t i = 0; i < m; ++i) {
            int32_t l = 0;
            for (int64_t p = H.ptr[i]; p < H.ptr[i + 1]; ++p) l = std::max(l, H.level[H.dep[p]] + 1);
            H.level[i] = l;
        }
        H.unknowns.resize(m);
        std::iota(H.unknowns.begin(), H.unknowns.end(), 0);
        upload_sweep(H, Y->fwd, device);
    }
    // ---- backward sweep: columns of L as rows of L^T, lead unknowns only
    {
        HostSweep H;
        H.n = m;
        H.ptr.assign(m + 1, 0);
        for (int64_t j = 0; j < n_lead; ++j) H.ptr[j + 1] = F.Lp[j + 1] - F.Lp[j] - 1;
        for (int64_t i = 0; i < m; ++i) H.ptr[i + 1] += H.ptr[i];
        H.dep.resize(H.ptr[m]); H.val.resize(H.ptr[m]);
        for (int64_t j = 0; j < n_lead; ++j) {
            int64_t q = H.ptr[j];
            for (int64_t p = F.Lp[j] + 1; p < F.Lp[j + 1]; ++p, ++q) { H.dep[q] = F.Li[p]; H.val[q] = F.Lx[p]; }
        }
        H.inv_diag.assign(m, 1.0);
        for (int64_t i = 0; i < n_lead; ++i) H.inv_diag[i] = 1.0 / F.Lx[F.Lp[i]];
        H.level.assign(m, 0);
        for (int64_t j = n_lead - 1; j >= 0; --j) {
            int32_t l = 0;
            for (int64_t p = H.ptr[j]; p < H.ptr[j + 1]; ++p) if (H.dep[p] < n_lead) l = std::max(l, H.level[H.dep[p]] + 1);
            H.level[j] = l;
        }
        H.unknowns.resize(n_lead);
        std::iota(H.unknowns.begin(), H.unknowns.end(), 0);
        upload_sweep(H, Y->bwd, device);
    }
    Y->z.alloc(std::max<int64_t>(m, 1));
    Y->x.alloc(std::max<int64_t>(m, 1));
    Y->flags.alloc(std::max<int64_t>(Y->fwd.levels + Y->bwd.levels, 1));

    // ---- dense tail: S = M22 - L21 L21^T, Cholesky, explicit inverse (all on the device)
    if (n_tail > 0) {
        int64_t tail_def = 0;
        build_dense_tail(C, F, n_lead, n_tail, Y->tail_inv, Y->tail_inv_t, &tail_def);
        Y->n_deficient += tail_def;
        Y->tail_tmp.alloc(n_tail);
    }
    Y->launches_per_solve = 1 + 2 + (n_tail > 0 ? 2 : 0);
    Y->alg_bytes = 2 * (12 * Y->nnz_L + 8 * m) + 24 * m;
    CUADMM_CUDA(cudaDeviceSynchronize());
    return Y.release();
}

}  // namespace cuadmm

using namespace cuadmm;

void cuadmm_ysolve_s::solve(const double* d_rhs_, double* d_y_, cudaStream_t st) {
    if (m == 0) return;
    const int nl = fwd.levels + bwd.levels;
    CUADMM_CUDA(cudaMemsetAsync(flags.p, 0, sizeof(int32_t) * (size_t)std::max(nl, 1), st));
    // forward: z = L11^-1 P rhs (lead), z_tail = P rhs - L21 z_lead
    launch_sweep(fwd, d_rhs_, perm.p, z.p, flags.p, nullptr, nullptr, done_flag, st);
    if (n_tail > 0) {
        const int blocks = (int)((n_tail + 7) / 8);
        // x_tail = L22^-T L22^-1 z_tail, scattered into y
        tail_gemv_kernel<<<blocks, 256, 0, st>>>(n_tail, tail_inv.p, z.p + n_lead, tail_tmp.p, 1, nullptr, nullptr, 0, done_flag);
        tail_gemv_kernel<<<blocks, 256, 0, st>>>(n_tail, tail_inv_t.p, tail_tmp.p, x.p + n_lead, 0, d_y_, perm.p, n_lead, done_flag);
        CUADMM_CUDA(cudaGetLastError());
    }
    // backward: x_lead = L11^-T (z_lead - L21^T x_tail), scattered into y
    launch_sweep(bwd, z.p, nullptr, x.p, flags.p + fwd.levels, d_y_, perm.p, done_flag, st);
}

extern "C" {

int cuadmm_ysolve_create(int64_t m, int64_t vec_len, int64_t nnz, const int32_t* h_A_rowptr, const int32_t* h_A_colind,
                         const double* h_A_val, double eps, int device, cuadmm_ysolve_t** out) {
    return guarded([&] {
        CUADMM_REQUIRE(out != nullptr, "out is null");
        *out = nullptr;
        *out = ysolve_create(m, vec_len, nnz, h_A_rowptr, h_A_colind, h_A_val, eps, device);
    });
}

void cuadmm_ysolve_destroy(cuadmm_ysolve_t* ys) { delete ys; }

int cuadmm_ysolve(cuadmm_ysolve_t* ys, const double* d_rhs, double* d_y, void* stream) {
    return guarded([&] {
        CUADMM_REQUIRE(ys && d_rhs && d_y, "null argument");
        DeviceGuard g(ys->device);
        ys->solve(d_rhs, d_y, (cudaStream_t)stream);
    });
}

int cuadmm_ysolve_host(cuadmm_ysolve_t* ys, const double* h_rhs, double* h_y) {
    return guarded([&] {
        CUADMM_REQUIRE(ys && h_rhs && h_y, "null argument");
        DeviceGuard g(ys->device);
        if (ys->d_rhs.n != ys->m) { ys->d_rhs.alloc(std::max<int64_t>(ys->m, 1)); ys->d_y.alloc(std::max<int64_t>(ys->m, 1)); }
        ys->d_rhs.upload(h_rhs, ys->m);
        ys->solve(ys->d_rhs.p, ys->d_y.p, 0);
        ys->d_y.download(h_y, ys->m);
        CUADMM_CUDA(cudaStreamSynchronize(0));
    });
}

int cuadmm_ysolve_stats(const cuadmm_ysolve_t* ys, int64_t out[8]) {
    return guarded([&] {
        CUADMM_REQUIRE(ys && out, "null argument");
        out[0] = ys->nnz_aat; out[1] = ys->nnz_L; out[2] = std::max(ys->fwd.levels, ys->bwd.levels);
        out[3] = ys->n_tail; out[4] = ys->launches_per_solve; out[5] = ys->alg_bytes;
        out[6] = ys->n_deficient; out[7] = ys->fwd.grid;
    });
}

int cuadmm_ysolve_perm(const cuadmm_ysolve_t* ys, int32_t* perm) {
    return guarded([&] {
        CUADMM_REQUIRE(ys && perm, "null argument");
        std::copy(ys->h_perm.begin(), ys->h_perm.end(), perm);
    });
}

}  // extern "C"
