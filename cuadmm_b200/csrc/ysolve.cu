// ysolve.cu — y = (A A^T + eps I)^-1 rhs on the device.
//
// Replaces the reference's per-iteration host path (src/solver.cu:487-500, 704-717):
//   perform_permutation -> cudaDeviceSynchronize -> D2H -> cholmod_solve2 (host, simplicial LDL^T)
//   -> H2D -> perform_permutation
// with: gather-permute fused into a level-scheduled sparse forward sweep, a dense GEMV pair on the
// explicitly inverted trailing block of the factor (where the elimination DAG degenerates into a
// chain), a sparse backward sweep with the scatter-permute fused in.  Nothing leaves the GPU and
// there is no host synchronisation.
#include "ysolve.h"
#include "dense.h"
#include <algorithm>
#include <numeric>
#include <stdlib.h>

namespace cuadmm {

static constexpr int kTriThreads = 256;
static constexpr int kLongRowNnz = 24;     // rows with more dependencies get a whole warp
static constexpr int kNarrowSlots = 160;   // levels with at most this many warp-slots run inside one CTA
static constexpr int kNarrowThreads = 1024;

// Work of a triangular sweep is cut into warp-slots: a slot is either 8 short rows (4 lanes each) or
// 1 long row (32 lanes), all of one dependency level.  The levels are grouped into phases:
//   wide level   -> one launch of tri_wide_kernel, one warp per slot, the stream order is the barrier;
//   narrow run   -> consecutive narrow levels in ONE CTA (tri_narrow_kernel), __syncthreads between
//                   levels (~0.1 us instead of a ~2.5 us kernel boundary or a multi-us global counter
//                   handshake; the deep end of the elimination DAG is hundreds of levels a few rows wide).
// No spin-waiting anywhere, so no co-residency requirement and nothing to deadlock.
// Unknown u:  x[u] = (rhs[u] - sum_p val[p] * x[dep[p]]) * inv_diag[u].
// Everything of a slot that does not depend on x (structure, values, right-hand side) is loaded by
// SlotWork::load and can therefore be issued one level AHEAD of the level barrier; finish() then
// only has the x loads on its critical path.  The deep end of the elimination DAG is latency-bound:
// this turns a chain of five dependent global loads per level into one.
struct SlotWork {
    static constexpr int K = 4;        // entries per lane kept in registers
    int32_t u; int is_long;
    double rv, invd;
    int64_t p, p1; int step;
    int32_t d[K]; double v[K];

    __device__ __forceinline__ void load(int64_t s, int lane, const int32_t* __restrict__ slot_rows,
            const int32_t* __restrict__ slot_info, const int64_t* __restrict__ ptr, const int32_t* __restrict__ dep,
            const double* __restrict__ val, const double* __restrict__ inv_diag, const double* __restrict__ rhs,
            const int32_t* __restrict__ rhs_gather) {
        is_long = slot_info[s] >> 30;
        u = slot_rows[8 * s + (is_long ? 0 : (lane >> 2))];
        step = is_long ? 32 : 4;
        p = 0; p1 = 0; rv = 0.0; invd = 0.0;
        if (u >= 0) {
            p = ptr[u] + (is_long ? lane : (lane & 3));
            p1 = ptr[u + 1];
            invd = inv_diag[u];
            rv = rhs_gather ? rhs[rhs_gather[u]] : rhs[u];
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int64_t q = p + (int64_t)k * step;
            if (q < p1) { d[k] = dep[q]; v[k] = val[q]; } else { d[k] = -1; v[k] = 0.0; }
        }
    }
    __device__ __forceinline__ void finish(int lane, const int32_t* __restrict__ dep, const double* __restrict__ val,
                                           double* x, double* out_scatter, const int32_t* __restrict__ out_perm) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) if (d[k] >= 0) acc = fma(v[k], x[d[k]], acc);
        for (int64_t q = p + (int64_t)K * step; q < p1; q += step) acc = fma(val[q], x[dep[q]], acc);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (is_long) {
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            acc += __shfl_xor_sync(0xffffffffu, acc, 8);
            acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        }
        const bool writer = (u >= 0) && (is_long ? (lane == 0) : ((lane & 3) == 0));
        if (writer) {
            const double r = (rv - acc) * invd;
            x[u] = r;
            if (out_scatter) out_scatter[out_perm[u]] = r;
        }
    }
};

#define CUADMM_TRI_PARAMS                                                                                   \
    const int32_t* __restrict__ slot_rows, const int32_t* __restrict__ slot_info, const int64_t* __restrict__ ptr, \
    const int32_t* __restrict__ dep, const double* __restrict__ val, const double* __restrict__ inv_diag,   \
    const double* __restrict__ rhs, const int32_t* __restrict__ rhs_gather, double* x, double* out_scatter, \
    const int32_t* __restrict__ out_perm, const int* __restrict__ done_flag
#define CUADMM_TRI_LOAD(W, S) (W).load((S), lane, slot_rows, slot_info, ptr, dep, val, inv_diag, rhs, rhs_gather)
#define CUADMM_TRI_FINISH(W) (W).finish(lane, dep, val, x, out_scatter, out_perm)

__global__ void __launch_bounds__(kTriThreads) tri_wide_kernel(int64_t slot0, int64_t slot1, CUADMM_TRI_PARAMS) {
    if (done_flag && *done_flag) return;
    const int lane = threadIdx.x & 31;
    const int64_t s = slot0 + (((int64_t)blockIdx.x * kTriThreads + threadIdx.x) >> 5);
    if (s >= slot1) return;
    SlotWork w;
    CUADMM_TRI_LOAD(w, s);
    CUADMM_TRI_FINISH(w);
}

// levels [level0, level1) of lvl_ptr inside one CTA of NW warps; the first slot of the next level is
// prefetched (static part) before the barrier of the current one
template <int NW>
__device__ __forceinline__ void tri_level_loop(const int64_t* __restrict__ lvl_ptr, int level0, int level1, int lane, int warp,
        CUADMM_TRI_PARAMS) {
    (void)done_flag;
    if (level0 >= level1) return;
    SlotWork w;
    int64_t s0 = lvl_ptr[level0];
    int64_t s1 = lvl_ptr[level0 + 1];
    bool have = (s0 + warp) < s1;
    if (have) CUADMM_TRI_LOAD(w, s0 + warp);
    for (int l = level0; l < level1; ++l) {
        if (have) {
            CUADMM_TRI_FINISH(w);
            for (int64_t s = s0 + warp + NW; s < s1; s += NW) { SlotWork t; CUADMM_TRI_LOAD(t, s); CUADMM_TRI_FINISH(t); }
        }
        s0 = s1;
        have = false;
        if (l + 1 < level1) {
            s1 = lvl_ptr[l + 2];
            have = (s0 + warp) < s1;
            if (have) CUADMM_TRI_LOAD(w, s0 + warp);
        }
        __syncthreads();   // also makes this CTA's global writes visible to its own later loads
    }
}

__global__ void __launch_bounds__(kNarrowThreads) tri_narrow_kernel(int level0, int level1, const int64_t* __restrict__ level_ptr,
                                                                    CUADMM_TRI_PARAMS) {
    if (done_flag && *done_flag) return;
    tri_level_loop<kNarrowThreads / 32>(level_ptr, level0, level1, threadIdx.x & 31, threadIdx.x >> 5,
        slot_rows, slot_info, ptr, dep, val, inv_diag, rhs, rhs_gather, x, out_scatter, out_perm, done_flag);
}

// Subtree parallelism: every CTA owns one subtree of the elimination tree (all of whose dependencies
// are inside the subtree, or already final), and walks the subtree's own levels with __syncthreads.
// Thousands of independent deep chains (one per block neighbourhood of a moment relaxation) thus cost
// ONE launch and depth x (one x-load latency) instead of depth x (kernel boundary).
__global__ void __launch_bounds__(kTriThreads) tri_subtree_kernel(const int64_t* __restrict__ sub_off,
        const int64_t* __restrict__ sub_lvl_ptr, CUADMM_TRI_PARAMS) {
    if (done_flag && *done_flag) return;
    const int64_t base = sub_off[blockIdx.x];
    const int nl = (int)(sub_off[blockIdx.x + 1] - base) - 1;
    tri_level_loop<kTriThreads / 32>(sub_lvl_ptr + base, 0, nl, threadIdx.x & 31, threadIdx.x >> 5,
        slot_rows, slot_info, ptr, dep, val, inv_diag, rhs, rhs_gather, x, out_scatter, out_perm, done_flag);
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

static void launch_subtrees(const TriSweep& S, const double* rhs, const int32_t* gather, double* x,
                            double* out_scatter, const int32_t* out_perm, const int* done, cudaStream_t st) {
    if (S.n_sub == 0) return;
    tri_subtree_kernel<<<(unsigned)S.n_sub, kTriThreads, 0, st>>>(S.sub_off.p, S.sub_lvl_ptr.p, S.slot_rows.p, S.slot_info.p,
        S.ptr.p, S.dep.p, S.val.p, S.inv_diag.p, rhs, gather, x, out_scatter, out_perm, done);
}

static int launch_sweep(const TriSweep& S, const double* rhs, const int32_t* gather, double* x,
                        double* out_scatter, const int32_t* out_perm, const int* done, cudaStream_t st) {
    int launches = 0;
    if (S.subtrees_first && S.n_sub) { launch_subtrees(S, rhs, gather, x, out_scatter, out_perm, done, st); ++launches; }
    for (const TriSweep::Phase& ph : S.phases) {
        if (ph.narrow) {
            tri_narrow_kernel<<<1, kNarrowThreads, 0, st>>>(ph.level0, ph.level1, S.level_ptr.p, S.slot_rows.p, S.slot_info.p,
                S.ptr.p, S.dep.p, S.val.p, S.inv_diag.p, rhs, gather, x, out_scatter, out_perm, done);
        } else {
            const int64_t s0 = S.h_level_ptr[ph.level0], s1 = S.h_level_ptr[ph.level1];
            const int grid = (int)(((s1 - s0) * 32 + kTriThreads - 1) / kTriThreads);
            tri_wide_kernel<<<grid, kTriThreads, 0, st>>>(s0, s1, S.slot_rows.p, S.slot_info.p, S.ptr.p, S.dep.p, S.val.p,
                S.inv_diag.p, rhs, gather, x, out_scatter, out_perm, done);
        }
        ++launches;
    }
    if (!S.subtrees_first && S.n_sub) { launch_subtrees(S, rhs, gather, x, out_scatter, out_perm, done, st); ++launches; }
    CUADMM_CUDA(cudaGetLastError());
    return launches;
}

// host: pull structure -> device sweep
struct HostSweep {
    std::vector<int64_t> ptr;
    std::vector<int32_t> dep;
    std::vector<double> val, inv_diag;
    int64_t n = 0;                  // unknown id space
    std::vector<int32_t> order;     // the unknowns of this sweep in a valid (topological) solve order
    std::vector<int32_t> sub;       // per unknown id: subtree id, or -1 = top part
    int64_t n_sub = 0;
    bool subtrees_first = true;     // forward: subtrees then top; backward: top then subtrees
};

// slots for a list of unknowns that all belong to one level
static int64_t emit_level_slots(const HostSweep& H, std::vector<int32_t>& rows, int level,
                                std::vector<int32_t>& slot_rows, std::vector<int32_t>& slot_info) {
    std::vector<int32_t> longs, shorts;
    for (int32_t u : rows) ((H.ptr[u + 1] - H.ptr[u]) > kLongRowNnz ? longs : shorts).push_back(u);
    std::stable_sort(longs.begin(), longs.end(), [&](int32_t a, int32_t b) {
        return H.ptr[a + 1] - H.ptr[a] > H.ptr[b + 1] - H.ptr[b]; });
    int64_t cnt = 0;
    for (int32_t u : longs) {
        slot_rows.push_back(u);
        for (int t = 1; t < 8; ++t) slot_rows.push_back(-1);
        slot_info.push_back(level | (1 << 30));
        ++cnt;
    }
    for (size_t t = 0; t < shorts.size(); t += 8) {
        for (size_t q = t; q < t + 8; ++q) slot_rows.push_back(q < shorts.size() ? shorts[q] : -1);
        slot_info.push_back(level);
        ++cnt;
    }
    return cnt;
}

static void upload_sweep(const HostSweep& H, TriSweep& S) {
    S.n_unknowns = (int64_t)H.order.size();
    S.nnz = (int64_t)H.dep.size();
    S.subtrees_first = H.subtrees_first;
    // levels inside each part (a subtree, or the top): dependencies that live in another part are
    // final by construction of the launch order and do not count
    std::vector<int32_t> level(H.n, 0);
    for (int32_t u : H.order) {
        int32_t l = 0;
        for (int64_t p = H.ptr[u]; p < H.ptr[u + 1]; ++p) {
            const int32_t d = H.dep[p];
            if (H.sub[d] == H.sub[u]) l = std::max(l, level[d] + 1);
        }
        level[u] = l;
    }
    std::vector<int32_t> slot_rows, slot_info;
    // ---- subtree part: slots grouped by (subtree, level); deepest subtrees first
    std::vector<std::vector<int32_t>> members(H.n_sub);
    std::vector<int32_t> depth(H.n_sub, 0);
    for (int32_t u : H.order) if (H.sub[u] >= 0) {
        members[H.sub[u]].push_back(u);
        depth[H.sub[u]] = std::max(depth[H.sub[u]], level[u] + 1);
    }
    std::vector<int32_t> sub_order(H.n_sub);
    std::iota(sub_order.begin(), sub_order.end(), 0);
    std::stable_sort(sub_order.begin(), sub_order.end(), [&](int32_t a, int32_t b) { return depth[a] > depth[b]; });
    std::vector<int64_t> sub_off(1, 0), sub_lvl_ptr;
    int max_sub_depth = 0;
    for (int32_t t : sub_order) {
        std::vector<std::vector<int32_t>> by_level(depth[t]);
        for (int32_t u : members[t]) by_level[level[u]].push_back(u);
        for (int l = 0; l < depth[t]; ++l) {
            sub_lvl_ptr.push_back((int64_t)slot_info.size());
            emit_level_slots(H, by_level[l], l, slot_rows, slot_info);
        }
        sub_lvl_ptr.push_back((int64_t)slot_info.size());
        sub_off.push_back((int64_t)sub_lvl_ptr.size());
        max_sub_depth = std::max(max_sub_depth, depth[t]);
    }
    S.n_sub = H.n_sub;
    S.sub_depth = max_sub_depth;
    // ---- top part: global levels, wide levels one launch each, narrow runs in one CTA
    int maxlev = -1;
    for (int32_t u : H.order) if (H.sub[u] < 0) maxlev = std::max(maxlev, (int)level[u]);
    S.levels = maxlev + 1;
    std::vector<std::vector<int32_t>> top_by_level(std::max(S.levels, 0));
    for (int32_t u : H.order) if (H.sub[u] < 0) top_by_level[level[u]].push_back(u);
    S.h_level_ptr.assign(S.levels + 1, (int64_t)slot_info.size());
    std::vector<int64_t> level_slots(std::max(S.levels, 1), 0);
    for (int l = 0; l < S.levels; ++l) {
        S.h_level_ptr[l] = (int64_t)slot_info.size();
        level_slots[l] = emit_level_slots(H, top_by_level[l], l, slot_rows, slot_info);
        S.h_level_ptr[l + 1] = (int64_t)slot_info.size();
    }
    S.n_slots = (int64_t)slot_info.size();
    int narrow_slots = kNarrowSlots;
    if (const char* e = getenv("CUADMM_SWEEP_NARROW_SLOTS")) narrow_slots = atoi(e);
    S.phases.clear();
    for (int l = 0; l < S.levels;) {
        if (level_slots[l] <= narrow_slots) {
            int e = l;
            while (e < S.levels && level_slots[e] <= narrow_slots) ++e;
            S.phases.push_back({true, l, e});
            l = e;
        } else {
            S.phases.push_back({false, l, l + 1});
            ++l;
        }
    }
    S.ptr.upload(H.ptr);
    std::vector<int32_t> dep(H.dep); if (dep.empty()) dep.push_back(0);
    std::vector<double> val(H.val); if (val.empty()) val.push_back(0.0);
    S.dep.upload(dep); S.val.upload(val);
    S.inv_diag.upload(H.inv_diag);
    if (slot_rows.empty()) { slot_rows.assign(8, -1); slot_info.push_back(0); }
    S.slot_rows.upload(slot_rows); S.slot_info.upload(slot_info);
    S.level_ptr.upload(S.h_level_ptr);
    if (sub_lvl_ptr.empty()) sub_lvl_ptr.push_back(0);
    S.sub_off.upload(sub_off); S.sub_lvl_ptr.upload(sub_lvl_ptr);
}

// Disjoint subtrees of the elimination tree (lead part only), each at most `cap` rows: v roots a
// subtree when its own subtree fits and its parent's does not (or the parent is in the dense tail).
static int64_t choose_subtrees(const CholFactor& F, int64_t n_lead, std::vector<int32_t>& sub) {
    const int64_t n = F.n;
    sub.assign(n, -1);
    int64_t cap = 32768;
    if (const char* e = getenv("CUADMM_SWEEP_SUBTREE_CAP")) cap = atoll(e);
    if (cap <= 0) return 0;
    int64_t min_size = 2;         // (measured: peeling small subtrees off into extra top levels costs more than their CTAs)
    if (const char* e = getenv("CUADMM_SWEEP_SUBTREE_MIN")) min_size = atoll(e);
    std::vector<int64_t> size(n, 1);
    for (int64_t v = 0; v < n_lead; ++v) {
        const int32_t p = F.parent[v];
        if (p >= 0 && p < n_lead) size[p] += size[v];
    }
    int64_t n_sub = 0;
    for (int64_t v = n_lead - 1; v >= 0; --v) {
        const int32_t p = F.parent[v];
        const bool parent_top = (p < 0 || p >= n_lead || sub[p] < 0);
        if (!parent_top) { sub[v] = sub[p]; continue; }
        const bool parent_fits = (p >= 0 && p < n_lead && size[p] <= cap);
        if (size[v] <= cap && !parent_fits && size[v] >= min_size) sub[v] = (int32_t)n_sub++;
    }
    return n_sub;
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
    int64_t max_tail = 10240;
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
        chol_symbolic(M, perm0, F, &C);
    }
    chol_numeric(C, F, n_lead);
    Y->n_lead = n_lead; Y->n_tail = n_tail;
    Y->nnz_L = F.nnz();
    Y->n_deficient = F.n_deficient;
    Y->h_perm = F.perm;
    Y->perm.upload(F.perm);

    std::vector<int32_t> sub;
    const int64_t n_sub = choose_subtrees(F, n_lead, sub);
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
        H.order.resize(m);
        std::iota(H.order.begin(), H.order.end(), 0);
        H.sub = sub; H.n_sub = n_sub; H.subtrees_first = true;
        upload_sweep(H, Y->fwd);
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
        H.order.resize(n_lead);
        for (int64_t j = 0; j < n_lead; ++j) H.order[j] = (int32_t)(n_lead - 1 - j);
        // tail unknowns are final before the backward sweep starts: give them their own "part"
        H.sub = sub;
        for (int64_t i = n_lead; i < m; ++i) H.sub[i] = -2;
        H.n_sub = n_sub; H.subtrees_first = false;
        upload_sweep(H, Y->bwd);
    }
    Y->z.alloc(std::max<int64_t>(m, 1));
    Y->x.alloc(std::max<int64_t>(m, 1));

    // ---- dense tail: S = M22 - L21 L21^T, Cholesky, explicit inverse (all on the device)
    if (n_tail > 0) {
        int64_t tail_def = 0;
        build_dense_tail(C, F, n_lead, n_tail, Y->tail_inv, Y->tail_inv_t, &tail_def);
        Y->n_deficient += tail_def;
        Y->tail_tmp.alloc(n_tail);
    }
    Y->launches_per_solve = (int)(Y->fwd.phases.size() + Y->bwd.phases.size()) + (n_tail > 0 ? 2 : 0) + (n_sub > 0 ? 2 : 0);
    Y->alg_bytes = 2 * (12 * Y->nnz_L + 8 * m) + 24 * m;
    CUADMM_CUDA(cudaDeviceSynchronize());
    return Y.release();
}

}  // namespace cuadmm

using namespace cuadmm;

void cuadmm_ysolve_s::solve(const double* d_rhs_, double* d_y_, cudaStream_t st) {
    if (m == 0) return;
    // forward: z = L11^-1 P rhs (lead), z_tail = P rhs - L21 z_lead
    launch_sweep(fwd, d_rhs_, perm.p, z.p, nullptr, nullptr, done_flag, st);
    if (n_tail > 0) {
        const int blocks = (int)((n_tail + 7) / 8);
        // x_tail = L22^-T L22^-1 z_tail, scattered into y
        tail_gemv_kernel<<<blocks, 256, 0, st>>>(n_tail, tail_inv.p, z.p + n_lead, tail_tmp.p, 1, nullptr, nullptr, 0, done_flag);
        tail_gemv_kernel<<<blocks, 256, 0, st>>>(n_tail, tail_inv_t.p, tail_tmp.p, x.p + n_lead, 0, d_y_, perm.p, n_lead, done_flag);
        CUADMM_CUDA(cudaGetLastError());
    }
    // backward: x_lead = L11^-T (z_lead - L21^T x_tail), scattered into y
    launch_sweep(bwd, z.p, nullptr, x.p, d_y_, perm.p, done_flag, st);
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
        out[6] = ys->n_deficient; out[7] = ys->fwd.n_sub * 100000 + ys->fwd.sub_depth;
    });
}

int cuadmm_ysolve_perm(const cuadmm_ysolve_t* ys, int32_t* perm) {
    return guarded([&] {
        CUADMM_REQUIRE(ys && perm, "null argument");
        std::copy(ys->h_perm.begin(), ys->h_perm.end(), perm);
    });
}

}  // extern "C"
