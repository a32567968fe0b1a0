// Block-level device code of the swarm hot path (sm_100a).  One CTA owns one env:
//   * the env's FP64 integrator state lives in shared memory for the whole kernel
//     (step, auto-reset burn-in and rasterise never round-trip through HBM);
//   * the O(N^2) pair forces run in FP32 on an FP32 hi/lo split of the positions staged in
//     shared memory.  For N <= 512 every UNORDERED pair is evaluated once (the pair force is
//     exactly antisymmetric): a warp owns a 32-locust tile, walks the other tiles with a
//     lane rotation, keeps its own forces in registers and hands the reaction forces round
//     the warp with shuffles; reactions are combined through fixed shared-memory slots so
//     the result is bitwise reproducible.  Larger N falls back to an ordered-pair loop with
//     T targets per thread in registers and broadcast LDS.128 sources;
//   * reward is a warp-shuffle + shared-memory reduction in FP64;
//   * the occupancy grid is a shared-memory-privatised histogram with warp-aggregated
//     atomics (match.any), written out as a streaming zero fill + sparse scatter.  Its
//     counter table aliases the force scratch (the two phases never overlap).
//
// Reference semantics restated here: fed_gym/envs/multiagent.py:30-115,
// fed_gym/agents/state_processors.py:25-42.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/swarm_b200.h"
#include "swarm_philox.cuh"

namespace swarm {

constexpr int kSymMaxLocusts = 512;     // unordered-pair path (MODE 1) up to here
constexpr unsigned kFull = 0xffffffffu;

// Kernel-side parameter block (derived from SwarmParams on the host).
// Pair geometry is done in coordinates pre-scaled by c = log2(e):  exp(-r) = 2^(-c r), and
// d/(r + eps) is scale free if eps is scaled too, so the scaling costs nothing per pair.
struct KP {
    int E, N, A, G, n_burn, max_steps;
    double sigma, wind, dt, half_w, y_hi, cscale;
    float F, nInvL, U, Gv, eps_s;
    uint2 key;
    uint32_t env_off;
};

struct Smem {
    double2* xs;      // N   locust positions (FP64 integrator state)
    double2* as;      // A   agent positions
    double2* act;     // A   current actions
    double2* an;      // A   current agent noise row (unscaled)
    double* red;      // 32  reduction scratch
    double* box;      // 2   rasteriser: mean x
    // ---- scratch, aliased between the force phase and the rasteriser
    float4* src;      // FP32 sources (hi_x, hi_y, lo_x, lo_y), x = hi + lo to ~2^-48.
                      //   MODE 1: nt tiles x 64 (each 32-tile stored twice: wrap-free lane+k reads) + A agents
                      //   MODE 2/4: N locusts + A agents
    float2* slot;     // MODE 1: nt x (1 + nt/2) x 32 reaction-force partial sums
    uint32_t* table;  // G*G packed cell counters: lo16 locusts, hi16 agents
    int* cid;         // N+A: cell written out by this point (or -1)
};

__host__ __device__ inline size_t smem_align(size_t v) { return (v + 15) & ~size_t(15); }
__host__ __device__ inline int n_tiles(int N) { return (N + 31) >> 5; }

__host__ __device__ inline size_t smem_src_bytes(int N, int A, bool sym) {
    return smem_align(sizeof(float4) * (sym ? (size_t)n_tiles(N) * 64 + A : (size_t)N + A));
}
__host__ __device__ inline size_t smem_fixed_bytes(int N, int A) {
    return smem_align(sizeof(double2) * N) + 3 * smem_align(sizeof(double2) * A) + smem_align(sizeof(double) * 32) +
           smem_align(sizeof(double) * 2);
}
__host__ __device__ inline size_t smem_raster_bytes(int N, int A, int G) {
    return smem_align(sizeof(uint32_t) * G * G) + smem_align(sizeof(int) * (N + A));
}
__host__ __device__ inline size_t smem_bytes(int N, int A, int G, bool raster, bool sym) {
    const int nt = n_tiles(N);
    size_t force = smem_src_bytes(N, A, sym) + (sym ? smem_align(sizeof(float2) * nt * (1 + nt / 2) * 32) : 0);
    size_t rast = raster ? smem_raster_bytes(N, A, G) : 0;
    return smem_fixed_bytes(N, A) + (force > rast ? force : rast);
}

__device__ __forceinline__ Smem carve(unsigned char* base, int N, int A, int G, bool sym) {
    Smem s;
    size_t o = 0;
    s.xs = reinterpret_cast<double2*>(base + o);  o += smem_align(sizeof(double2) * N);
    s.as = reinterpret_cast<double2*>(base + o);  o += smem_align(sizeof(double2) * A);
    s.act = reinterpret_cast<double2*>(base + o); o += smem_align(sizeof(double2) * A);
    s.an = reinterpret_cast<double2*>(base + o);  o += smem_align(sizeof(double2) * A);
    s.red = reinterpret_cast<double*>(base + o);  o += smem_align(sizeof(double) * 32);
    s.box = reinterpret_cast<double*>(base + o);  o += smem_align(sizeof(double) * 2);
    s.src = reinterpret_cast<float4*>(base + o);
    s.slot = reinterpret_cast<float2*>(base + o + smem_src_bytes(N, A, sym));
    s.table = reinterpret_cast<uint32_t*>(base + o);
    s.cid = reinterpret_cast<int*>(base + o + smem_align(sizeof(uint32_t) * G * G));
    return s;
}

// ------------------------------------------------------------------------------------------
// multiagent.py:70-86  x_update = cutoff; x += dt*v + noise; cutoff   (FP64, no FMA contraction
// so that an injected v reproduces the reference's three roundings exactly)
__device__ __forceinline__ void move_particle(double2& p, double2 v, double2 n, double dt, double sigma) {
    if (p.y <= 0.0) {
        p.y = 0.0;
        v.x = 0.0;
        if (v.y <= 0.0) v.y = 0.0;
    }
    p.x = __dadd_rn(p.x, __dadd_rn(__dmul_rn(dt, v.x), __dmul_rn(sigma, n.x)));
    p.y = __dadd_rn(p.y, __dadd_rn(__dmul_rn(dt, v.y), __dmul_rn(sigma, n.y)));
    if (p.y <= 0.0) p.y = 0.0;
}

// The threads that cooperate on one env (today: the whole CTA).
struct Grp {
    int tid;   // thread index inside the group
    int n;     // threads in the group (multiple of 32)
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
// Pair weight, multiagent.py:65-68,100-113:  w = s(r)/(r+eps), s(r) = F exp(-r/L) - exp(-r),
// in log2(e)-scaled coordinates:  w = (F 2^(-r/L) - 2^(-r)) / (r + c eps).
// Positions arrive as FP32 hi/lo pairs of the FP64 state, so the difference
//   d = (hi_i - hi_j) + (lo_i - lo_j)
// carries ~2^-24 RELATIVE error however close the two particles are (plain FP32 positions lose
// the direction of close pairs, where s/(r+eps) is steepest).  Coincident points (and the self
// pair of the ordered loop) have d = 0 exactly and contribute exactly 0.
template <bool PRECISE>
__device__ __forceinline__ float pair_weight(const float dx, const float dy, const KP& kp) {
    if (PRECISE) {
        const float r = __fsqrt_rn(__fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
        const float s = __fmaf_rn(kp.F, exp2f(__fmul_rn(r, kp.nInvL)), -exp2f(-r));
        return __fdiv_rn(s, __fadd_rn(r, kp.eps_s));
    } else {
        // r2 >= 1e-30 keeps rsqrt finite for coincident points (their dx,dy are 0 anyway)
        const float r2 = fmaf(dx, dx, fmaf(dy, dy, 1e-30f));
        float rinv, e1, e2, inv;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(r2));
        const float r = r2 * rinv;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-r));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(r * kp.nInvL));
        const float s = fmaf(kp.F, e2, -e1);
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(r + kp.eps_s));
        return s * inv;
    }
}

// ordered pair: force of source q on target tg
template <bool PRECISE>
__device__ __forceinline__ void pair_ordered(const float4 q, const float4 tg, const KP& kp, float& ax, float& ay) {
    const float dx = (q.x - tg.x) + (q.z - tg.z);
    const float dy = (q.y - tg.y) + (q.w - tg.w);
    const float w = pair_weight<PRECISE>(dx, dy, kp);
    ax = fmaf(w, dx, ax);
    ay = fmaf(w, dy, ay);
}

__device__ __forceinline__ float4 split_hilo(const double2 p, const double c) {
    const double sx = p.x * c, sy = p.y * c;
    const float hx = (float)sx, hy = (float)sy;
    return make_float4(hx, hy, (float)(sx - (double)hx), (float)(sy - (double)hy));
}

// Stage the scaled FP32 hi/lo sources: locusts then agents.
template <int MODE>
__device__ __forceinline__ void stage_sources(const Smem& sm, const KP& kp, const Grp& g) {
    const int N = kp.N, A = kp.A;
    if (MODE == 1) {
        const int nt = n_tiles(N);
        // pad lanes sit far away: as sources they contribute exactly 0 (both exponentials underflow)
        const float4 pad = make_float4(1e15f, 0.f, 0.f, 0.f);
        for (int j = g.tid; j < nt * 32; j += g.n) {
            const float4 q = j < N ? split_hilo(sm.xs[j], kp.cscale) : pad;
            float4* t = sm.src + (j >> 5) * 64 + (j & 31);
            t[0] = q;
            t[32] = q;
        }
        for (int k = g.tid; k < A; k += g.n) sm.src[nt * 64 + k] = split_hilo(sm.as[k], kp.cscale);
    } else {
        for (int i = g.tid; i < N; i += g.n) sm.src[i] = split_hilo(sm.xs[i], kp.cscale);
        for (int k = g.tid; k < A; k += g.n) sm.src[N + k] = split_hilo(sm.as[k], kp.cscale);
    }
}

// Rotation steps K0..K1 of one tile pair: in step k lane l meets element (l+k)%32 of the other
// tile (read wrap-free from the doubled tile).  The own force accumulates in (ax,ay); the
// reaction on the met element accumulates in (bx,by), which moves one lane down per step so
// that it follows its element.  On return lane l holds the reaction for element (l+K1)%32.
template <int K0, int K1, bool PRECISE>
__device__ __forceinline__ void tile_sym(const float4* __restrict__ tl, const float4 tg, const KP& kp, const int nxt,
                                         float& ax, float& ay, float& bx, float& by) {
    bx = 0.f;
    by = 0.f;
#pragma unroll
    for (int k = K0; k <= K1; ++k) {
        const float4 q = tl[k];
        const float dx = (q.x - tg.x) + (q.z - tg.z);
        const float dy = (q.y - tg.y) + (q.w - tg.w);
        const float w = pair_weight<PRECISE>(dx, dy, kp);
        ax = fmaf(w, dx, ax);
        ay = fmaf(w, dy, ay);
        bx = fmaf(-w, dx, bx);
        by = fmaf(-w, dy, by);
        if (k < K1) {
            bx = __shfl_sync(kFull, bx, nxt);
            by = __shfl_sync(kFull, by, nxt);
        }
    }
}

// MODE 1: all locust-locust pairs once.  Thread = locust j (tile I = warp, lane).  Tile I meets
// tiles I+1..I+floor((nt-1)/2) fully, tile I+nt/2 (nt even) half each way, and itself.
template <bool PRECISE>
__device__ __forceinline__ void forces_sym(const Smem& sm, const KP& kp, const Grp& g, float& ax, float& ay) {
    const int lane = g.tid & 31, I = g.tid >> 5;
    const int nt = n_tiles(kp.N), nslots = 1 + nt / 2;
    const int nxt = (lane + 1) & 31;
    const float4* S = sm.src;
    const float4 tg = S[I * 64 + lane];
    float bx, by;
    ax = 0.f;
    ay = 0.f;
    // own tile: offsets 1..15 both ways, offset 16 one way (each such pair appears in two lanes), offset 0 = self
    tile_sym<1, 15, PRECISE>(S + I * 64 + lane, tg, kp, nxt, ax, ay, bx, by);
    sm.slot[(I * nslots) * 32 + ((lane + 15) & 31)] = make_float2(bx, by);
    pair_ordered<PRECISE>(S[I * 64 + lane + 16], tg, kp, ax, ay);
    const int nfull = (nt - 1) >> 1;
    for (int o = 1; o <= nfull; ++o) {
        int B = I + o;
        if (B >= nt) B -= nt;
        tile_sym<0, 31, PRECISE>(S + B * 64 + lane, tg, kp, nxt, ax, ay, bx, by);
        sm.slot[(B * nslots + o) * 32 + ((lane + 31) & 31)] = make_float2(bx, by);
    }
    if ((nt & 1) == 0) {
        // lane offsets 0..15 from the lower tile, 16..31 (= 1..16 seen from the partner) from the upper
        const int o = nt >> 1;
        const int B = I < o ? I + o : I - o;
        const int koff = I < o ? 0 : 1;
        tile_sym<0, 15, PRECISE>(S + B * 64 + lane + koff, tg, kp, nxt, ax, ay, bx, by);
        sm.slot[(B * nslots + o) * 32 + ((lane + 15 + koff) & 31)] = make_float2(bx, by);
    }
    g.sync();
    for (int o = 0; o < nslots; ++o) {          // fixed order: bitwise reproducible
        const float2 r = sm.slot[(I * nslots + o) * 32 + lane];
        ax += r.x;
        ay += r.y;
    }
    const float4* ag = S + nt * 64;             // agents act on locusts only (multiagent.py:108-113)
    for (int k = 0; k < kp.A; ++k) pair_ordered<PRECISE>(ag[k], tg, kp, ax, ay);
}

// MODE 2/4: ordered pairs, T targets per thread (j = tid + t*blockDim.x), broadcast LDS.128 sources.
template <int T, bool PRECISE>
__device__ __forceinline__ void forces_ordered(const Smem& sm, const KP& kp, const Grp& g, float (&vx)[T], float (&vy)[T]) {
    const int N = kp.N, S = kp.N + kp.A;
    float4 tg[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int j = g.tid + t * g.n;
        tg[t] = sm.src[j < N ? j : N - 1];
        vx[t] = 0.f;
        vy[t] = 0.f;
    }
#pragma unroll 4
    for (int i = 0; i < S; ++i) {
        const float4 q = sm.src[i];
#pragma unroll
        for (int t = 0; t < T; ++t) pair_ordered<PRECISE>(q, tg[t], kp, vx[t], vy[t]);
    }
}

template <int MODE>
struct ModeT { static constexpr int T = MODE == 1 ? 1 : MODE; };

// Forces on this thread's targets + block-wide reward = -mean_j |v_j|^2 (pre-cutoff v, wind and
// gravity added).  Needs the staged sources visible; ends after a barrier.
template <int MODE, bool PRECISE>
__device__ __forceinline__ double pair_forces(const Smem& sm, const KP& kp, const Grp& g,
                                              float (&vx)[ModeT<MODE>::T], float (&vy)[ModeT<MODE>::T]) {
    constexpr int T = ModeT<MODE>::T;
    if (MODE == 1) forces_sym<PRECISE>(sm, kp, g, vx[0], vy[0]);
    else forces_ordered<T, PRECISE>(sm, kp, g, vx, vy);
    double e = 0.0;
#pragma unroll
    for (int t = 0; t < T; ++t) {
        vx[t] += kp.U;
        vy[t] += kp.Gv;
        if (g.tid + t * g.n < kp.N) e += (double)vx[t] * (double)vx[t] + (double)vy[t] * (double)vy[t];
    }
    e = warp_sum(e);
    if ((g.tid & 31) == 0) sm.red[g.tid >> 5] = e;
    g.sync();
    double tot = 0.0;
    const int nw = g.n >> 5;
    for (int w = 0; w < nw; ++w) tot += sm.red[w];   // same order in every thread
    return -tot / (double)kp.N;
}

// SwarmEnv._step on the shared-memory state.  Preconditions: sm.xs/as/act/an filled and
// visible (a __syncthreads since their last write); nx[t] = unscaled noise of own target t.
// Postcondition: state updated and visible to the whole block; force scratch free again.
template <int MODE, bool PRECISE>
__device__ __forceinline__ double env_step(const Smem& sm, const KP& kp, const Grp& g,
                                           const double2 (&nx)[ModeT<MODE>::T], float* v_out) {
    constexpr int T = ModeT<MODE>::T;
    // multiagent.py:33-38  agents move first: v_action (+wind on x) through x_update
    for (int k = g.tid; k < kp.A; k += g.n) {
        double2 a = sm.as[k];
        double2 w = sm.act[k];
        w.x = __dadd_rn(w.x, kp.wind);
        move_particle(a, w, sm.an[k], kp.dt, kp.sigma);
        sm.as[k] = a;
    }
    g.sync();
    stage_sources<MODE>(sm, kp, g);      // old x, NEW xa (multiagent.py:39)
    g.sync();
    float vx[T], vy[T];
    const double reward = pair_forces<MODE, PRECISE>(sm, kp, g, vx, vy);
    // multiagent.py:40  locusts move with the pre-cutoff v just computed
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int j = g.tid + t * g.n;
        if (j < kp.N) {
            if (v_out) reinterpret_cast<float2*>(v_out)[j] = make_float2(vx[t], vy[t]);
            double2 p = sm.xs[j];
            move_particle(p, make_double2((double)vx[t], (double)vy[t]), nx[t], kp.dt, kp.sigma);
            sm.xs[j] = p;
        }
    }
    g.sync();
    return reward;
}

// SwarmEnv._reset (multiagent.py:46-63) for env e: draws (injected or Philox), n_burn burn-in
// steps with noise row k, then the frozen row n_burn is stored for all later steps (Q1).
template <int MODE, bool PRECISE>
__device__ __noinline__ void env_reset(const Smem& sm, const KP& kp, const Grp& g, int e, uint32_t episode,
                                       const bool inj, const SwarmInjectedDraws& dr, const SwarmState& st) {
    constexpr int T = ModeT<MODE>::T;
    const int N = kp.N, A = kp.A;
    DrawCtx ctx;
    ctx.key = kp.key;
    ctx.env = kp.env_off + (uint32_t)e;
    ctx.episode = episode;
    for (int i = g.tid; i < N; i += g.n)
        sm.xs[i] = inj ? reinterpret_cast<const double2*>(dr.x0)[(size_t)e * N + i]
                       : draw_uniform2(ctx, STREAM_X0, i);
    for (int k = g.tid; k < A; k += g.n)
        sm.as[k] = inj ? reinterpret_cast<const double2*>(dr.xa0)[(size_t)e * A + k]
                       : draw_uniform2(ctx, STREAM_XA0, k);
    const int rows = kp.n_burn + 1;
    for (int r = 0; r <= kp.n_burn; ++r) {
        double2 nx[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int j = g.tid + t * g.n;
            nx[t] = make_double2(0.0, 0.0);
            if (j < N)
                nx[t] = inj ? reinterpret_cast<const double2*>(dr.particle_noise)[((size_t)e * rows + r) * N + j]
                            : draw_normal2(ctx, STREAM_NOISE_X, r, j);
        }
        if (r == kp.n_burn) {   // frozen row: kept in HBM for every later step
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const int j = g.tid + t * g.n;
                if (j < N) reinterpret_cast<double2*>(st.noise_x)[(size_t)e * N + j] = nx[t];
            }
            for (int k = g.tid; k < A; k += g.n)
                reinterpret_cast<double2*>(st.noise_a)[(size_t)e * A + k] =
                    inj ? reinterpret_cast<const double2*>(dr.agent_noise)[((size_t)e * rows + r) * A + k]
                        : draw_normal2(ctx, STREAM_NOISE_A, r, k);
            break;
        }
        for (int k = g.tid; k < A; k += g.n) {
            sm.act[k] = inj ? reinterpret_cast<const double2*>(dr.burn_actions)[((size_t)e * kp.n_burn + r) * A + k]
                            : draw_normal2(ctx, STREAM_BURN, r, k);
            sm.an[k] = inj ? reinterpret_cast<const double2*>(dr.agent_noise)[((size_t)e * rows + r) * A + k]
                           : draw_normal2(ctx, STREAM_NOISE_A, r, k);
        }
        g.sync();
        env_step<MODE, PRECISE>(sm, kp, g, nx, nullptr);
    }
    g.sync();
}

// ------------------------------------------------------------------------------------------
// np.searchsorted(edges, p, side='right') for edges = linspace(lo, hi, G+1) as numpy builds
// them: e[i] = fl(fl(i*step) + lo) for i < G, e[G] = hi  (numpy/_core/function_base.py).
__device__ __forceinline__ double edge_at(int i, double lo, double hi, double step, int G) {
    return i >= G ? hi : __dadd_rn(__dmul_rn((double)i, step), lo);
}

// The bin is guessed with a reciprocal multiply and then corrected against the exact edges.
__device__ __forceinline__ int count_le(double p, double lo, double hi, double step, double inv_step, int G) {
    if (!(p == p)) return G + 1;                 // NaN sorts last
    const double q = (p - lo) * inv_step;
    int g = q < -1.0 ? -1 : (q > (double)G ? G : (int)floor(q));
    while (g < G && edge_at(g + 1, lo, hi, step, G) <= p) ++g;
    while (g >= 0 && edge_at(g, lo, hi, step, G) > p) --g;
    return g + 1;                                // #{i in [0,G] : e[i] <= p}
}

// Streaming zero fill of one env's (G,G,2) f32 grid by n threads (evict-first stores).
__device__ __forceinline__ void raster_zero_fill(float* __restrict__ grid_e, int cells, int tid, int n) {
    if ((cells & 1) == 0) {
        float4* g4 = reinterpret_cast<float4*>(grid_e);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = tid; i < cells / 2; i += n) __stcs(g4 + i, z);
    } else {
        float2* g2 = reinterpret_cast<float2*>(grid_e);
        for (int i = tid; i < cells; i += n) __stcs(g2 + i, make_float2(0.f, 0.f));
    }
}

// SwarmStateProcessor.process_state (state_processors.py:25-42) of the positions xs (N) / as (A) in
// shared memory, by the thread group g.  grid_e: (G,G,2) f32, pos_e: (A,2) u8 of this env.
// The grid must have been zero-filled (raster_zero_fill) before a barrier that precedes this call.
// Uses sm.table / sm.cid / sm.box (the force scratch must be dead unless they do not alias).
__device__ __forceinline__ void env_raster(const Smem& sm, const double2* __restrict__ xs, const double2* __restrict__ as,
                                           const KP& kp, const Grp& g, float* __restrict__ grid_e,
                                           uint8_t* __restrict__ pos_e) {
    const int N = kp.N, A = kp.A, G = kp.G, P = N + A, cells = G * G;
    // phase 0: one thread walks the sequential FP64 mean (np.mean(vstack([x,xa]),axis=0)[0] is a
    // plain left-to-right sum, ~23 cycles per dependent DADD); everyone else clears the counters.
    if (g.tid == 0) {
        double s = 0.0;
        for (int i = 0; i < N; ++i) s = __dadd_rn(s, xs[i].x);
        for (int k = 0; k < A; ++k) s = __dadd_rn(s, as[k].x);
        sm.box[0] = s / (double)P;
    } else {
        for (int i = g.tid - 1; i < cells; i += g.n - 1) sm.table[i] = 0u;
    }
    g.sync();
    // phase 1: bin every point in FP64 against numpy's edges, count with warp-aggregated atomics
    const double m = sm.box[0];
    const double lo_x = m - kp.half_w, hi_x = m + kp.half_w;
    const double step_x = (hi_x - lo_x) / (double)G;
    const double lo_y = 0.0, hi_y = kp.y_hi;
    const double step_y = (hi_y - lo_y) / (double)G;
    const double inv_x = 1.0 / step_x, inv_y = 1.0 / step_y;
    for (int base = 0; base < P; base += g.n) {
        const int p = base + g.tid;
        uint32_t key = 0xffffffffu;
        int cell = 0;
        if (p < P) {
            const bool agent = p >= N;
            const double2 q = agent ? as[p - N] : xs[p];
            const int cx = count_le(q.x, lo_x, hi_x, step_x, inv_x, G);
            const int cy = count_le(q.y, lo_y, hi_y, step_y, inv_y, G);
            if (agent) {   // np.digitize -> bin+1, clamped to G-1 (state_processors.py:35-40)
                pos_e[2 * (p - N) + 0] = (uint8_t)(cx < G - 1 ? cx : G - 1);
                pos_e[2 * (p - N) + 1] = (uint8_t)(cy < G - 1 ? cy : G - 1);
            }
            const int bx = (q.x == hi_x) ? cx - 2 : cx - 1;   // histogramdd right-edge fix-up
            const int by = (q.y == hi_y) ? cy - 2 : cy - 1;
            if (bx >= 0 && bx < G && by >= 0 && by < G) {
                cell = bx * G + by;
                key = ((uint32_t)cell << 1) | (agent ? 1u : 0u);
            }
        }
        const uint32_t peers = __match_any_sync(kFull, key);
        int mine = -1;
        if (key != 0xffffffffu && (__ffs(peers) - 1) == (g.tid & 31)) {
            const uint32_t inc = (uint32_t)__popc(peers) << ((key & 1u) ? 16 : 0);
            const uint32_t old = atomicAdd(&sm.table[cell], inc);
            if (old == 0u) mine = cell;               // first arrival writes the cell out
        }
        if (p < P) sm.cid[p] = mine;
    }
    g.sync();
    // phase 2: sparse scatter of the non-zero cells over the zero fill
    float2* g2 = reinterpret_cast<float2*>(grid_e);
    for (int p = g.tid; p < P; p += g.n) {
        const int c = sm.cid[p];
        if (c >= 0) {
            const uint32_t w = sm.table[c];
            g2[c] = make_float2(__fdiv_rn((float)(w & 0xffffu), (float)N),
                                A > 0 ? __fdiv_rn((float)(w >> 16), (float)A) : 0.f);
        }
    }
}

}  // namespace swarm
