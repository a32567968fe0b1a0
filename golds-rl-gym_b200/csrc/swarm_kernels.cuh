// Block-level device code of the swarm hot path (sm_100a).  A CTA works on one env at a time
// with two warp-specialised thread groups: the FORCE group (dynamics, reward, auto-reset) and
// the RASTER group (occupancy grid of the env the force group finished last), so that the
// rasteriser's latency chains and HBM stores run under the next env's force phase:
//   * an env's FP64 integrator state, its frozen noise row and its actions are prefetched
//     (cp.async) into one of two shared-memory STAGE buffers while the previous env is being
//     stepped, and stay there for the whole step (step, auto-reset burn-in and rasterise
//     never round-trip through HBM);
//   * the O(N^2) pair forces run in FP32 on an FP32 hi/lo split of the positions staged in
//     shared memory.  For N <= 512 every UNORDERED pair is evaluated once (the pair force is
//     exactly antisymmetric): a warp owns a 32-locust tile, walks the other tiles with a
//     lane rotation, keeps its own forces in registers and hands the reaction forces round
//     the warp with shuffles; reactions are combined through fixed shared-memory slots so
//     the result is bitwise reproducible.  Larger N falls back to an ordered-pair loop with
//     T targets per thread in registers and broadcast LDS.128 sources;
//   * reward is a warp-shuffle + shared-memory reduction in FP64;
//   * the occupancy grid is a shared-memory-privatised histogram with warp-aggregated
//     atomics (match.any), written out as a streaming zero fill + sparse scatter; the
//     counter table is cleaned cell by cell by the threads that scatter, never re-cleared.
//
// Everything here is __forceinline__ on purpose: the shared-memory pointers of `Smem` must
// stay in registers with a known address space (a spilled pointer turns every LDS/STS into
// a generic LD/ST).
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

// One STAGE buffer = everything the step of one env reads from HBM:
//   x (N) | noise_x (N) | xa (A) | noise_a (A) | raw actions (A x 16 B) | misc (16 B: elapsed)
struct Stage {
    double2* xs;            // N   locust positions (FP64 integrator state, updated in place)
    double2* nx;            // N   locust noise row of the current step (unscaled)
    double2* as;            // A   agent positions
    double2* an;            // A   agent noise row
    unsigned char* araw;    // A x 16 B: the env's actions as they sit in HBM (f32 pair or f64 pair)
    int* misc;              // [0] = TimeLimit elapsed steps
};

struct Smem {
    Stage st;         // the stage buffer of the env being stepped
    double2* act;     // A   actions after conversion / clipping
    double* red;      // 32  reduction scratch
    double* box;      // 2   rasteriser: mean x
    // ---- force scratch
    float4* src;      // FP32 sources (hi_x, hi_y, lo_x, lo_y), x = hi + lo to ~2^-48.
                      //   MODE 1/3: nt 32-tiles x 64 (each tile stored twice: wrap-free lane+k reads) + A agents
                      //   MODE 2/4: N locusts + A agents
    float2* slot;     // MODE 1/3: nt x nslots x 32 reaction-force partial sums
    // ---- rasteriser (present when the kernel rasterises)
    double2* rx[2];   // 2 x (N+A): post-step positions handed from the force to the raster group
    uint32_t* table;  // G*G packed cell counters: lo16 locusts, hi16 agents
    int* cid;         // N+A: cell written out by this point (or -1)
};

__host__ __device__ inline size_t smem_align(size_t v) { return (v + 15) & ~size_t(15); }
__host__ __device__ inline int n_tiles(int N) { return (N + 31) >> 5; }

__host__ __device__ inline size_t smem_stage_bytes(int N, int A) { return 16 * ((size_t)2 * N + 3 * A + 1); }
__host__ __device__ inline int sym_tiles(int N, int sym);
__host__ __device__ inline size_t smem_src_bytes(int N, int A, int sym) {
    return smem_align(sizeof(float4) * (sym ? (size_t)sym_tiles(N, sym) * 64 + A : (size_t)N + A));
}
__host__ __device__ inline size_t smem_fixed_bytes(int A) {
    return smem_align(sizeof(double2) * A) + smem_align(sizeof(double) * 32) + smem_align(sizeof(double) * 2);
}
// sym: 0 = ordered pairs (MODE 2/4), 1 = unordered, 32-wide tiles (MODE 1), 2 = unordered, 64-wide (MODE 3)
__host__ __device__ inline int sym_tiles(int N, int sym) { return sym == 2 ? 2 * ((N + 63) >> 6) : n_tiles(N); }
__host__ __device__ inline int sym_slots(int N, int sym) {
    return sym == 2 ? 1 + ((N + 63) >> 6) / 2 : 1 + n_tiles(N) / 2;
}
__host__ __device__ inline size_t smem_force_bytes(int N, int A, int sym) {
    return smem_src_bytes(N, A, sym) +
           (sym ? smem_align(sizeof(float2) * sym_tiles(N, sym) * sym_slots(N, sym) * 32) : 0);
}
__host__ __device__ inline size_t smem_raster_bytes(int N, int A, int G) {
    return 2 * smem_align(sizeof(double2) * (N + A)) + smem_align(sizeof(uint32_t) * G * G) +
           smem_align(sizeof(int) * (N + A));
}
// n_stage: stage buffers (2 in the pipelined step kernel, 1 in reset/forces, 0 in the rasteriser)
__host__ __device__ inline size_t smem_bytes(int N, int A, int G, int n_stage, bool force, bool raster, int sym) {
    return n_stage * smem_stage_bytes(N, A) + smem_fixed_bytes(A) + (force ? smem_force_bytes(N, A, sym) : 0) +
           (raster ? smem_raster_bytes(N, A, G) : 0);
}

__device__ __forceinline__ Stage stage_at(unsigned char* base, int N, int A, int b) {
    unsigned char* p = base + (size_t)b * smem_stage_bytes(N, A);
    Stage s;
    s.xs = reinterpret_cast<double2*>(p);
    s.nx = s.xs + N;
    s.as = s.nx + N;
    s.an = s.as + A;
    s.araw = reinterpret_cast<unsigned char*>(s.an + A);
    s.misc = reinterpret_cast<int*>(s.araw + 16 * (size_t)A);
    return s;
}

__device__ __forceinline__ Smem carve(unsigned char* base, int N, int A, int G, int n_stage, bool force, int sym) {
    Smem s;
    s.st = stage_at(base, N, A, 0);
    size_t o = (size_t)n_stage * smem_stage_bytes(N, A);
    s.act = reinterpret_cast<double2*>(base + o); o += smem_align(sizeof(double2) * A);
    s.red = reinterpret_cast<double*>(base + o);  o += smem_align(sizeof(double) * 32);
    s.box = reinterpret_cast<double*>(base + o);  o += smem_align(sizeof(double) * 2);
    s.src = reinterpret_cast<float4*>(base + o);
    s.slot = reinterpret_cast<float2*>(base + o + smem_src_bytes(N, A, sym));
    if (force) o += smem_force_bytes(N, A, sym);
    s.rx[0] = reinterpret_cast<double2*>(base + o); o += smem_align(sizeof(double2) * (N + A));
    s.rx[1] = reinterpret_cast<double2*>(base + o); o += smem_align(sizeof(double2) * (N + A));
    s.table = reinterpret_cast<uint32_t*>(base + o); o += smem_align(sizeof(uint32_t) * G * G);
    s.cid = reinterpret_cast<int*>(base + o);
    return s;
}

// ------------------------------------------------------------------------------------------
// cp.async (LDGSTS): HBM -> shared memory without staging through registers
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gmem_src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

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

// Named barriers (bar.sync id, n) of the warp-specialised step kernel.  0 stays __syncthreads.
enum : int { BAR_FORCE = 1, BAR_RASTER = 2, BAR_FULL0 = 3, BAR_FULL1 = 4, BAR_EMPTY0 = 5, BAR_EMPTY1 = 6 };

template <int ID>
__device__ __forceinline__ void bar_sync(int n) { asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(n) : "memory"); }
template <int ID>
__device__ __forceinline__ void bar_arrive(int n) { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "r"(n) : "memory"); }

// A group of whole warps that cooperates on one env: tid in [0,n), n a multiple of 32.
template <int BAR>
struct GrpT {
    int tid;
    int n;
    __device__ __forceinline__ void sync() const { bar_sync<BAR>(n); }
};
typedef GrpT<BAR_FORCE> Grp;     // dynamics
typedef GrpT<BAR_RASTER> RGrp;   // rasteriser

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

// MODE: 1 = unordered pairs, a warp owns a 32-locust tile, 1 target per lane
//       3 = unordered pairs, a warp owns a 64-locust super-tile, 2 targets per lane (half the
//           shared-memory loads and shuffles per pair; used when N pads to 64 as well as to 32)
//       2/4 = ordered pairs, 2/4 targets per thread (N > 512)
template <int MODE>
struct ModeT {
    static constexpr int T = MODE == 1 ? 1 : (MODE == 3 ? 2 : MODE);
    static constexpr int SYM = MODE == 1 ? 1 : (MODE == 3 ? 2 : 0);
};

// locust owned by thread g.tid as its t-th target
template <int MODE>
__device__ __forceinline__ int target_index(const Grp& g, int t) {
    return MODE == 3 ? ((g.tid >> 5) * 64 + t * 32 + (g.tid & 31)) : g.tid + t * g.n;
}

template <int MODE>
__device__ __forceinline__ float4* agent_sources(const Smem& sm, const KP& kp) {
    return sm.src + (ModeT<MODE>::SYM ? sym_tiles(kp.N, ModeT<MODE>::SYM) * 64 : kp.N);
}

// Stage the scaled FP32 hi/lo sources of the locusts.  Thread j stages element j.
template <int MODE>
__device__ __forceinline__ void stage_locusts(const Smem& sm, const KP& kp, const Grp& g) {
    const int N = kp.N;
    if (ModeT<MODE>::SYM) {
        const int nt = sym_tiles(N, ModeT<MODE>::SYM);
        // pad lanes sit far away: as sources they contribute exactly 0 (both exponentials underflow)
        const float4 pad = make_float4(1e15f, 0.f, 0.f, 0.f);
        for (int j = g.tid; j < nt * 32; j += g.n) {
            const float4 q = j < N ? split_hilo(sm.st.xs[j], kp.cscale) : pad;
            float4* t = sm.src + (j >> 5) * 64 + (j & 31);
            t[0] = q;
            t[32] = q;
        }
    } else {
        for (int i = g.tid; i < N; i += g.n) sm.src[i] = split_hilo(sm.st.xs[i], kp.cscale);
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
#pragma unroll 8
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

// MODE 1, part 1: all locust-locust pairs once.  Thread = locust j (tile I = warp, lane).  Tile I
// meets tiles I+1..I+floor((nt-1)/2) fully, tile I+nt/2 (nt even) half each way, and itself.
// The reaction sums land in sm.slot; the caller must barrier before forces_sym_finish.
template <bool PRECISE>
__device__ __forceinline__ void forces_sym_tiles(const Smem& sm, const KP& kp, const Grp& g, float& ax, float& ay) {
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
}

// MODE 1, part 2: add the reactions (fixed order: bitwise reproducible) and the agents' pull.
template <bool PRECISE>
__device__ __forceinline__ void forces_sym_finish(const Smem& sm, const KP& kp, const Grp& g, float& ax, float& ay) {
    const int lane = g.tid & 31, I = g.tid >> 5;
    const int nt = n_tiles(kp.N), nslots = 1 + nt / 2;
    for (int o = 0; o < nslots; ++o) {
        const float2 r = sm.slot[(I * nslots + o) * 32 + lane];
        ax += r.x;
        ay += r.y;
    }
    const float4 tg = sm.src[I * 64 + lane];
    const float4* ag = sm.src + nt * 64;        // agents act on locusts only (multiagent.py:108-113)
    for (int k = 0; k < kp.A; ++k) pair_ordered<PRECISE>(ag[k], tg, kp, ax, ay);
}

// ---- MODE 3: 64-wide super-tiles, two targets (A: first half, B: second half) per lane -------
template <bool REACT, bool PRECISE>
__device__ __forceinline__ void pair_sym(const float4 q, const float4 tg, const KP& kp, float& ax, float& ay,
                                         float& bx, float& by) {
    const float dx = (q.x - tg.x) + (q.z - tg.z);
    const float dy = (q.y - tg.y) + (q.w - tg.w);
    const float w = pair_weight<PRECISE>(dx, dy, kp);
    ax = fmaf(w, dx, ax);
    ay = fmaf(w, dy, ay);
    if (REACT) {
        bx = fmaf(-w, dx, bx);
        by = fmaf(-w, dy, by);
    }
}

// Rotation steps K0..K1 against one doubled half-tile tl (= S + H*64 + lane): in step k the lane
// meets element (lane+k)%32 and evaluates it against its A and/or B target; (bx,by) is the
// reaction on the met element and moves one lane down per step.  CONT: (bx,by) continues from an
// earlier call (shuffle before the first step too).  On return lane l holds element (l+K1)%32's.
template <int K0, int K1, bool DO_A, bool DO_B, bool REACT_A, bool REACT_B, bool CONT, bool PRECISE>
__device__ __forceinline__ void tile2(const float4* __restrict__ tl, const float4 tgA, const float4 tgB, const KP& kp,
                                      const int nxt, float& aAx, float& aAy, float& aBx, float& aBy, float& bx,
                                      float& by) {
#pragma unroll 8
    for (int k = K0; k <= K1; ++k) {
        if ((REACT_A || REACT_B) && (CONT || k > K0)) {
            bx = __shfl_sync(kFull, bx, nxt);
            by = __shfl_sync(kFull, by, nxt);
        }
        const float4 q = tl[k];
        if (DO_A) pair_sym<REACT_A, PRECISE>(q, tgA, kp, aAx, aAy, bx, by);
        if (DO_B) pair_sym<REACT_B, PRECISE>(q, tgB, kp, aBx, aBy, bx, by);
    }
}

// MODE 3, part 1.  Warp I owns super-tile I = half-tiles 2I (A targets) and 2I+1 (B targets).
//   own super-tile: A x A and B x B by offsets 1..15 both ways + 16 one way; A x B split by offset:
//                   (B target, A source) for offsets 0..15, (A target, B source) for 1..16;
//   super-tiles I+1 .. I+floor((nt2-1)/2): all 64 x 64 pairs;  I+nt2/2 (nt2 even): half each way.
template <bool PRECISE>
__device__ __forceinline__ void forces_sym64_tiles(const Smem& sm, const KP& kp, const Grp& g, float (&vx)[2],
                                                   float (&vy)[2]) {
    const int lane = g.tid & 31, I = g.tid >> 5;
    const int nt2 = (kp.N + 63) >> 6, nslots = 1 + nt2 / 2;
    const int nxt = (lane + 1) & 31;
    const float4* S = sm.src;
    const float4* tlA = S + (2 * I) * 64 + lane;
    const float4* tlB = tlA + 64;
    const float4 tgA = tlA[0], tgB = tlB[0];
    float aAx = 0.f, aAy = 0.f, aBx = 0.f, aBy = 0.f, bx = 0.f, by = 0.f;
    // sources = own A half
    tile2<0, 0, false, true, false, true, false, PRECISE>(tlA, tgA, tgB, kp, nxt, aAx, aAy, aBx, aBy, bx, by);
    tile2<1, 15, true, true, true, true, true, PRECISE>(tlA, tgA, tgB, kp, nxt, aAx, aAy, aBx, aBy, bx, by);
    sm.slot[((2 * I) * nslots) * 32 + ((lane + 15) & 31)] = make_float2(bx, by);
    tile2<16, 16, true, false, false, false, false, PRECISE>(tlA, tgA, tgB, kp, nxt, aAx, aAy, aBx, aBy, bx, by);
    // sources = own B half
    bx = 0.f; by = 0.f;
    tile2<1, 15, true, true, true, true, false, PRECISE>(tlB, tgA, tgB, kp, nxt, aAx, aAy, aBx, aBy, bx, by);
    tile2<16, 16, true, true, true, false, true, PRECISE>(tlB, tgA, tgB, kp, nxt, aAx, aAy, aBx, aBy, bx, by);
    sm.slot[((2 * I + 1) * nslots) * 32 + ((lane + 16) & 31)] = make_float2(bx, by);
    const int nfull = (nt2 - 1) >> 1;
    for (int o = 1; o <= nfull; ++o) {
        int J = I + o;
        if (J >= nt2) J -= nt2;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const int H = 2 * J + h;
            bx = 0.f; by = 0.f;
            tile2<0, 31, true, true, true, true, false, PRECISE>(S + H * 64 + lane, tgA, tgB, kp, nxt, aAx, aAy, aBx, aBy,
                                                                 bx, by);
            sm.slot[(H * nslots + o) * 32 + ((lane + 31) & 31)] = make_float2(bx, by);
        }
    }
    if ((nt2 & 1) == 0) {
        // lane offsets 0..15 from the lower super-tile, 16..31 (= 1..16 seen from the partner) from the upper
        const int o = nt2 >> 1;
        const int J = I < o ? I + o : I - o;
        const int koff = I < o ? 0 : 1;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const int H = 2 * J + h;
            bx = 0.f; by = 0.f;
            tile2<0, 15, true, true, true, true, false, PRECISE>(S + H * 64 + lane + koff, tgA, tgB, kp, nxt, aAx, aAy, aBx,
                                                                 aBy, bx, by);
            sm.slot[(H * nslots + o) * 32 + ((lane + 15 + koff) & 31)] = make_float2(bx, by);
        }
    }
    vx[0] = aAx; vy[0] = aAy; vx[1] = aBx; vy[1] = aBy;
}

// MODE 3, part 2: reactions (fixed order: bitwise reproducible) and the agents' pull.
template <bool PRECISE>
__device__ __forceinline__ void forces_sym64_finish(const Smem& sm, const KP& kp, const Grp& g, float (&vx)[2],
                                                    float (&vy)[2]) {
    const int lane = g.tid & 31, I = g.tid >> 5;
    const int nt2 = (kp.N + 63) >> 6, nslots = 1 + nt2 / 2;
    for (int o = 0; o < nslots; ++o) {
        const float2 ra = sm.slot[((2 * I) * nslots + o) * 32 + lane];
        const float2 rb = sm.slot[((2 * I + 1) * nslots + o) * 32 + lane];
        vx[0] += ra.x; vy[0] += ra.y;
        vx[1] += rb.x; vy[1] += rb.y;
    }
    const float4 tgA = sm.src[(2 * I) * 64 + lane], tgB = sm.src[(2 * I + 1) * 64 + lane];
    const float4* ag = sm.src + nt2 * 128;      // agents act on locusts only (multiagent.py:108-113)
    for (int k = 0; k < kp.A; ++k) {
        const float4 q = ag[k];
        pair_ordered<PRECISE>(q, tgA, kp, vx[0], vy[0]);
        pair_ordered<PRECISE>(q, tgB, kp, vx[1], vy[1]);
    }
}

// MODE 2/4: ordered pairs, T targets per thread (j = tid + t*n), broadcast LDS.128 sources.
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

// SwarmEnv.v_calculate (multiagent.py:88-115) on staged sources: v of this thread's targets (wind
// and gravity added, before any cutoff) and its share of sum_j |v_j|^2.  Needs the staged sources
// visible (a barrier since stage_*); contains one group barrier in MODE 1.
template <int MODE, bool PRECISE>
__device__ __forceinline__ double pair_forces(const Smem& sm, const KP& kp, const Grp& g,
                                              float (&vx)[ModeT<MODE>::T], float (&vy)[ModeT<MODE>::T]) {
    constexpr int T = ModeT<MODE>::T;
    if constexpr (MODE == 1) {
        forces_sym_tiles<PRECISE>(sm, kp, g, vx[0], vy[0]);
        g.sync();
        forces_sym_finish<PRECISE>(sm, kp, g, vx[0], vy[0]);
    } else if constexpr (MODE == 3) {
        forces_sym64_tiles<PRECISE>(sm, kp, g, vx, vy);
        g.sync();
        forces_sym64_finish<PRECISE>(sm, kp, g, vx, vy);
    } else {
        forces_ordered<T, PRECISE>(sm, kp, g, vx, vy);
    }
    double e = 0.0;
#pragma unroll
    for (int t = 0; t < T; ++t) {
        vx[t] += kp.U;
        vy[t] += kp.Gv;
        if (target_index<MODE>(g, t) < kp.N) e += (double)vx[t] * (double)vx[t] + (double)vy[t] * (double)vy[t];
    }
    return e;
}

// reward = -mean_j |v_j|^2 from the per-thread shares: warp shuffle, then one slot per warp.  The
// caller barriers between energy_put and energy_get.
__device__ __forceinline__ void energy_put(const Smem& sm, const Grp& g, double e) {
    e = warp_sum(e);
    if ((g.tid & 31) == 0) sm.red[g.tid >> 5] = e;
}
__device__ __forceinline__ double energy_get(const Smem& sm, const KP& kp, const Grp& g) {
    double tot = 0.0;
    const int nw = g.n >> 5;
    for (int w = 0; w < nw; ++w) tot += sm.red[w];   // same order in every thread
    return -tot / (double)kp.N;
}

// SwarmEnv._step on the stage buffer sm.st.  Preconditions: st.xs/nx/as/an and sm.act filled, each
// element written by the thread that owns it here (element i <-> thread i mod n) or visible through
// a barrier.  Postcondition: state updated and visible to the whole group; returns the reward.
template <int MODE, bool PRECISE>
__device__ __forceinline__ double env_step(const Smem& sm, const KP& kp, const Grp& g, float* v_out) {
    constexpr int T = ModeT<MODE>::T;
    // multiagent.py:33-38  agents move first: v_action (+wind on x) through x_update; the mover
    // stages the agent's NEW position as a force source (multiagent.py:39: old x, new xa)
    float4* ag = agent_sources<MODE>(sm, kp);
    for (int k = g.tid; k < kp.A; k += g.n) {
        double2 a = sm.st.as[k];
        double2 w = sm.act[k];
        w.x = __dadd_rn(w.x, kp.wind);
        move_particle(a, w, sm.st.an[k], kp.dt, kp.sigma);
        sm.st.as[k] = a;
        ag[k] = split_hilo(a, kp.cscale);
    }
    stage_locusts<MODE>(sm, kp, g);
    g.sync();
    float vx[T], vy[T];
    const double e = pair_forces<MODE, PRECISE>(sm, kp, g, vx, vy);
    energy_put(sm, g, e);
    // multiagent.py:40  locusts move with the pre-cutoff v just computed
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int j = target_index<MODE>(g, t);
        if (j < kp.N) {
            if (v_out) reinterpret_cast<float2*>(v_out)[j] = make_float2(vx[t], vy[t]);
            double2 p = sm.st.xs[j];
            move_particle(p, make_double2((double)vx[t], (double)vy[t]), sm.st.nx[j], kp.dt, kp.sigma);
            sm.st.xs[j] = p;
        }
    }
    g.sync();
    return energy_get(sm, kp, g);
}

// SwarmEnv._reset (multiagent.py:46-63) for env e on the stage buffer: draws (injected or Philox),
// n_burn burn-in steps with noise row k, then the frozen row n_burn is stored for all later
// steps (Q1).  Ends with the state visible to the whole group.
template <int MODE, bool PRECISE>
__device__ __forceinline__ void env_reset(const Smem& sm, const KP& kp, const Grp& g, int e, uint32_t episode,
                                          const bool inj, const SwarmInjectedDraws& dr, const SwarmState& st) {
    const int N = kp.N, A = kp.A;
    DrawCtx ctx;
    ctx.key = kp.key;
    ctx.env = kp.env_off + (uint32_t)e;
    ctx.episode = episode;
    g.sync();      // the previous step's readers of red / src are done
    for (int i = g.tid; i < N; i += g.n)
        sm.st.xs[i] = inj ? reinterpret_cast<const double2*>(dr.x0)[(size_t)e * N + i]
                          : draw_uniform2(ctx, STREAM_X0, i);
    for (int k = g.tid; k < A; k += g.n)
        sm.st.as[k] = inj ? reinterpret_cast<const double2*>(dr.xa0)[(size_t)e * A + k]
                          : draw_uniform2(ctx, STREAM_XA0, k);
    const int rows = kp.n_burn + 1;
    for (int r = 0; r <= kp.n_burn; ++r) {
        for (int j = g.tid; j < N; j += g.n) {
            const double2 z = inj ? reinterpret_cast<const double2*>(dr.particle_noise)[((size_t)e * rows + r) * N + j]
                                  : draw_normal2(ctx, STREAM_NOISE_X, r, j);
            sm.st.nx[j] = z;
            if (r == kp.n_burn) reinterpret_cast<double2*>(st.noise_x)[(size_t)e * N + j] = z;   // frozen row -> HBM
        }
        for (int k = g.tid; k < A; k += g.n) {
            const double2 z = inj ? reinterpret_cast<const double2*>(dr.agent_noise)[((size_t)e * rows + r) * A + k]
                                  : draw_normal2(ctx, STREAM_NOISE_A, r, k);
            sm.st.an[k] = z;
            if (r == kp.n_burn) {
                reinterpret_cast<double2*>(st.noise_a)[(size_t)e * A + k] = z;
            } else {
                sm.act[k] = inj ? reinterpret_cast<const double2*>(dr.burn_actions)[((size_t)e * kp.n_burn + r) * A + k]
                                : draw_normal2(ctx, STREAM_BURN, r, k);
            }
        }
        if (r == kp.n_burn) break;
        env_step<MODE, PRECISE>(sm, kp, g, nullptr);
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

// One-time clear of the counter table (afterwards env_raster leaves it clean).
__device__ __forceinline__ void raster_table_clear(const Smem& sm, int cells, const RGrp& g) {
    for (int i = g.tid; i < cells; i += g.n) sm.table[i] = 0u;
}

// SwarmStateProcessor.process_state (state_processors.py:25-42) of the points pts = [N locusts; A
// agents] in shared memory, by the thread group g.  grid_e: (G,G,2) f32, pos_e: (A,2) u8 of this env.
// Preconditions: grid_e zero-filled by this group (raster_zero_fill), sm.table all zero and pts
// visible (a barrier since).  Postcondition: sm.table all zero again, after a group barrier.
__device__ __forceinline__ void env_raster(const Smem& sm, const double2* __restrict__ pts, const KP& kp, const RGrp& g,
                                           float* __restrict__ grid_e, uint8_t* __restrict__ pos_e) {
    const int N = kp.N, A = kp.A, G = kp.G, P = N + A;
    // phase 0: one thread walks the sequential FP64 mean (np.mean(vstack([x,xa]),axis=0)[0] is a
    // plain left-to-right sum, ~23 cycles per dependent DADD)
    if (g.tid == 0) {
        double s = 0.0;
#pragma unroll 8
        for (int i = 0; i < P; ++i) s = __dadd_rn(s, pts[i].x);
        sm.box[0] = s / (double)P;
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
            const double2 q = pts[p];
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
    // phase 2: sparse scatter of the non-zero cells over the zero fill; the writer cleans its counter
    float2* g2 = reinterpret_cast<float2*>(grid_e);
    for (int p = g.tid; p < P; p += g.n) {
        const int c = sm.cid[p];
        if (c >= 0) {
            const uint32_t w = sm.table[c];
            sm.table[c] = 0u;
            g2[c] = make_float2(__fdiv_rn((float)(w & 0xffffu), (float)N),
                                A > 0 ? __fdiv_rn((float)(w >> 16), (float)A) : 0.f);
        }
    }
    g.sync();
}

}  // namespace swarm
