// ddz_kernels.cu -- sm_100a kernels and the C-ABI launchers of libddz_b200.so (see include/ddz_b200.h).
//
// Kernels (one CTA = kEnvs consecutive envs, thread t owns env b0+t for the rule work, the whole CTA
// cooperates on the stores):
//   k_reset          deal from host-supplied permutations                      (reference: CEnv.prepare)
//   k_transition     [apply the chosen move] [re-deal finished envs] count the legal moves of the new state
//   k_emit           CSR offsets (block scan + prefix of the per-CTA totals), move enumeration into
//                    actions_u64, then thermometer rows for actions_f32 and face as 128-bit coalesced stores
//   k_encode_actions packed moves -> [n,15,4] float32
// One env-step of a rollout = k_transition + k_emit (ddz_rollout_step).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/ddz_b200.h"
#include "ddz_device.cuh"

namespace ddz {

constexpr int kEnvs = 128;     // envs per CTA == threads per CTA
constexpr int kWarps = kEnvs / 32;

struct Workspace {
    int32_t* counts;   // [B]
    int32_t* blk;      // [nblk] legal moves per CTA
};
static inline int nblocks(int B) { return (B + kEnvs - 1) / kEnvs; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline Workspace ws_of(void* p, int B) {
    Workspace w; w.counts = (int32_t*)p; w.blk = (int32_t*)((char*)p + align_up((size_t)B * 4, 256));
    return w;
}

// ------------------------------------------------------------------------------------------------
// block helpers
// ------------------------------------------------------------------------------------------------
DDZ_DEV int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
DDZ_DEV long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
// exclusive scan of one int per thread over the CTA; *total = CTA sum.  smem: int[kWarps]
DDZ_DEV int block_exclusive_scan(int v, int* smem, int* total) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) smem[w] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < kWarps; i++) { int s = smem[i]; if (i < w) base += s; tot += s; }
    __syncthreads();
    *total = tot;
    return base + inc - v;
}
DDZ_DEV void stat_add(int64_t* stats, int slot, long long v) {
    v = warp_sum_ll(v);
    if ((threadIdx.x & 31) == 0 && v != 0 && stats) atomicAdd((unsigned long long*)&stats[slot], (unsigned long long)v);
}

// ------------------------------------------------------------------------------------------------
// k_reset
// ------------------------------------------------------------------------------------------------
DDZ_DEV bool redeal(Env& e, const int8_t* perm, const int8_t* lord_pile, int pool_games, int B, int b) {
    uint32_t games = e.meta >> 8;
    size_t row = (size_t)(games % (uint32_t)pool_games) * B + b;
    bool ok = deal(e, perm + 54 * row, lord_pile ? lord_pile[row] : 0);
    if (!ok) {  // refuse: leave an empty, finished env with the sticky error bit
#pragma unroll
        for (int q = 0; q < 3; q++) { e.hand[q] = e.hist[q] = e.recent[q] = 0; }
        e.meta = (e.meta & 0xFFFFFF00u) | 1u | 4u | 0x20u;
    }
    return ok;
}

__global__ void __launch_bounds__(kEnvs) k_reset(void* state, const int8_t* __restrict__ perm,
                                                 const int8_t* __restrict__ lord_pile, int pool_games,
                                                 int only_done, int64_t* stats, int B) {
    int b = blockIdx.x * kEnvs + threadIdx.x;
    int err = 0;
    if (b < B) {
        StateView v = view_of(state, B);
        Env e = load_env(v, b);
        if (!only_done || e.done()) {
            err = !redeal(e, perm, lord_pile, pool_games, B, b);
            store_env(v, b, e);
        }
    }
    stat_add(stats, 7, err);
}

// ------------------------------------------------------------------------------------------------
// k_transition: [step] [re-deal] count
// ------------------------------------------------------------------------------------------------
struct StepArgs {
    const int32_t* offsets; const uint64_t* actions; const void* choice; int mode;
    uint64_t seed, env0; uint32_t stepno; int32_t rewards[3];
    const int8_t* perm; const int8_t* lord_pile; int pool_games;
    int8_t* r; uint8_t* done; int8_t* cat; float* reward;
};

template <bool STEP, bool RAW>
__global__ void __launch_bounds__(kEnvs) k_transition(void* state, const uint64_t* __restrict__ raw_hands,
                                                      const uint64_t* __restrict__ raw_lasts, StepArgs a,
                                                      Workspace ws, int64_t* stats, int B) {
    __shared__ int s_red[kWarps];
    int b = blockIdx.x * kEnvs + threadIdx.x;
    bool valid = b < B;
    int n = 0;
    long long d_games = 0, d_lord = 0, d_down = 0, d_up = 0, d_steps = 0, d_retl = 0, d_retf = 0, d_err = 0, d_pass = 0;
    if (RAW) {
        if (valid) n = count_legal(masks_of(raw_hands[b]), raw_lasts[b]);
    } else if (valid) {
        StateView v = view_of(state, B);
        Env e = load_env(v, b);
        if (STEP) {
            int o_r = 0, o_cat = -1;
            float rw0 = 0.f, rw1 = 0.f, rw2 = 0.f;
            if (!e.done()) {
                int base = a.offsets[b], cnt = a.offsets[b + 1] - base;
                long long idx = -1;
                if (a.mode == DDZ_CHOICE_INDEX) idx = ((const int32_t*)a.choice)[b];
                else if (a.mode == DDZ_CHOICE_MOD) idx = cnt > 0 ? (long long)(((const uint32_t*)a.choice)[b] % (uint32_t)cnt) : -1;
                else if (a.mode == DDZ_CHOICE_PHILOX) idx = cnt > 0 ? (long long)(philox(a.seed, a.env0 + (uint64_t)b, a.stepno) % (uint32_t)cnt) : -1;
                else {
                    uint64_t want = ((const uint64_t*)a.choice)[b];
                    for (int i = 0; i < cnt; i++) if (a.actions[base + i] == want) { idx = i; break; }
                }
                if (idx < 0 || idx >= cnt) { e.meta |= 0x20u; d_err = 1; }
                else {
                    StepOut o = apply_move(e, a.actions[base + idx], a.rewards);
                    o_r = o.r; o_cat = o.cat; rw0 = o.reward[0]; rw1 = o.reward[1]; rw2 = o.reward[2];
                    d_steps = 1; d_pass = o.pass;
                    if (o.done) {
                        d_games = 1; d_lord = (o.winner == 1); d_down = (o.winner == 2); d_up = (o.winner == 0);
                        d_retl = (long long)rw1; d_retf = (long long)rw0 + (long long)rw2;
                    }
                }
            }
            if (a.r) a.r[b] = (int8_t)o_r;
            if (a.done) a.done[b] = (uint8_t)e.done();
            if (a.cat) a.cat[b] = (int8_t)o_cat;
            if (a.reward) { a.reward[3 * (size_t)b] = rw0; a.reward[3 * (size_t)b + 1] = rw1; a.reward[3 * (size_t)b + 2] = rw2; }
            if (a.perm && e.done()) d_err += !redeal(e, a.perm, a.lord_pile, a.pool_games, B, b);
            store_env(v, b, e);
        }
        if (!e.done()) n = count_legal(masks_of(hand_to_move(e)), trick_of(e));
    }
    if (ws.counts) {
        if (valid) ws.counts[b] = n;
        int w = warp_sum(n);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = w;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int i = 0; i < kWarps; i++) t += s_red[i]; ws.blk[blockIdx.x] = t; }
    }
    if (STEP) {
        stat_add(stats, 0, d_games); stat_add(stats, 1, d_lord); stat_add(stats, 2, d_down); stat_add(stats, 3, d_up);
        stat_add(stats, 4, d_steps); stat_add(stats, 5, d_retl); stat_add(stats, 6, d_retf); stat_add(stats, 7, d_err);
        stat_add(stats, 9, d_pass);
    }
}

// ------------------------------------------------------------------------------------------------
// k_emit: offsets, packed moves, thermometer rows
// ------------------------------------------------------------------------------------------------
DDZ_DEV float4 thermo_row(uint64_t packed, int rank, float s) {
    uint32_t c = (uint32_t)(packed >> (4 * rank)) & 15u;
    return make_float4(c > 0 ? s : 0.f, c > 1 ? s : 0.f, c > 2 ? s : 0.f, c > 3 ? s : 0.f);
}
template <int V> struct FaceCfg;
template <> struct FaceCfg<0> { static constexpr int C = 4; };
template <> struct FaceCfg<1> { static constexpr int C = 7; };
template <> struct FaceCfg<2> { static constexpr int C = 9; };
template <> struct FaceCfg<3> { static constexpr int C = 6; };

// the C count planes of env e for variant V (envi.py:87-217); the last two are the unknown-card plane,
// scaled by p[0], p[1] = share of the next / next-next player's hand (get_state_prob)
template <int V>
DDZ_DEV void face_planes(const Env& e, uint64_t* planes /*[C]*/, float* p /*[2]*/) {
    int s = e.cur(), prev = (s + 2) % 3, next = (s + 1) % 3;
    uint64_t hand = pick3(s, e.hand[0], e.hand[1], e.hand[2]);
    uint64_t taken = e.hist[0] + e.hist[1] + e.hist[2];
    int n = 0;
    planes[n++] = hand;
    planes[n++] = taken;
    if (V == 1 || V == 2) {
        planes[n++] = pick3(prev, e.hist[0], e.hist[1], e.hist[2]);
        planes[n++] = pick3(s, e.hist[0], e.hist[1], e.hist[2]);
        planes[n++] = pick3(next, e.hist[0], e.hist[1], e.hist[2]);
    }
    if (V == 2 || V == 3) {
        planes[n++] = pick3(prev, e.recent[0], e.recent[1], e.recent[2]);
        planes[n++] = pick3(next, e.recent[0], e.recent[1], e.recent[2]);   // (role-2)%3 == next
    }
    uint64_t unknown = kDeckPacked - taken - hand;
    planes[n++] = unknown;
    planes[n++] = unknown;
    int size1 = card_count(pick3(next, e.hand[0], e.hand[1], e.hand[2]));
    int size2 = card_count(pick3(prev, e.hand[0], e.hand[1], e.hand[2]));
    int tot = size1 + size2;
    p[0] = tot > 0 ? __fdiv_rn((float)size1, (float)tot) : 0.f;
    p[1] = tot > 0 ? __fdiv_rn((float)size2, (float)tot) : 0.f;
}

// the CTA's face block [nenv][C][15] float4 is contiguous: thread i writes vector i, i+kEnvs, ...
template <int C>
DDZ_DEV void store_face_rows(float4* __restrict__ dst, int nenv, const uint64_t* s_planes, const float* s_p) {
    const int nvec = nenv * C * 15;
    for (int i = threadIdx.x; i < nvec; i += kEnvs) {
        int row = i / 15, rank = i - row * 15;
        int env = row / C, c = row - env * C;
        float s = (c >= C - 2) ? s_p[env * 2 + (c - (C - 2))] : 1.f;
        dst[i] = thermo_row(s_planes[row], rank, s);
    }
}

template <int V>
__global__ void __launch_bounds__(kEnvs) k_face(const void* state, float4* __restrict__ face, int B) {
    constexpr int C = FaceCfg<V>::C;
    __shared__ uint64_t s_planes[kEnvs * C];
    __shared__ float s_p[kEnvs * 2];
    const int tid = threadIdx.x, b0 = blockIdx.x * kEnvs, b = b0 + tid;
    if (b < B) {
        Env e = load_env(view_of(const_cast<void*>(state), B), b);
        uint64_t pl[C]; float p[2];
        face_planes<V>(e, pl, p);
#pragma unroll
        for (int c = 0; c < C; c++) s_planes[tid * C + c] = pl[c];
        s_p[tid * 2] = p[0]; s_p[tid * 2 + 1] = p[1];
    }
    __syncthreads();
    store_face_rows<C>(face + (size_t)b0 * C * 15, min(kEnvs, B - b0), s_planes, s_p);
}

struct Emitter {
    uint64_t* out; long long pos, end;
    DDZ_DEV void operator()(uint64_t mv) { if (pos < end) out[pos] = mv; pos++; }
};

// V < 0: no face.  RAW: hands/lasts arrays instead of a state.
template <int V, bool RAW>
__global__ void __launch_bounds__(kEnvs) k_emit(const void* state, const uint64_t* __restrict__ raw_hands,
                                                const uint64_t* __restrict__ raw_lasts, Workspace ws,
                                                int32_t* __restrict__ offsets, uint64_t* actions_u64,
                                                float4* __restrict__ actions_f32, long long cap,
                                                float4* __restrict__ face, int64_t* stats, int B) {
    constexpr int C = FaceCfg<(V < 0 ? 0 : V)>::C;
    __shared__ int s_scan[kWarps];
    __shared__ long long s_base;
    __shared__ uint64_t s_planes[V < 0 ? 1 : kEnvs * C];
    __shared__ float s_p[V < 0 ? 1 : kEnvs * 2];
    const int tid = threadIdx.x, b0 = blockIdx.x * kEnvs, b = b0 + tid;
    const bool valid = b < B;
    const int nenv = min(kEnvs, B - b0);

    // ---- global offset of this CTA = sum of the totals of the CTAs before it
    long long part = 0;
    for (int i = tid; i < (int)blockIdx.x; i += kEnvs) part += ws.blk[i];
    part = warp_sum_ll(part);
    if (tid == 0) s_base = 0;
    __syncthreads();
    if ((tid & 31) == 0 && part) atomicAdd((unsigned long long*)&s_base, (unsigned long long)part);
    __syncthreads();
    const long long base = s_base;

    const int n = valid ? ws.counts[b] : 0;
    int total;
    const int local = block_exclusive_scan(n, s_scan, &total);
    const long long off = base + local;
    if (valid) offsets[b] = (int32_t)off;
    if (b == B - 1) {
        offsets[B] = (int32_t)(off + n);
        if (stats) {
            atomicAdd((unsigned long long*)&stats[8], (unsigned long long)(off + n));
            if (off + n > cap) atomicAdd((unsigned long long*)&stats[7], 1ull);
        }
    }

    // ---- rule work: one env per thread
    Env e;
    if (valid) {
        uint64_t hand, last;
        if (RAW) { hand = raw_hands[b]; last = raw_lasts[b]; }
        else {
            StateView v = view_of(const_cast<void*>(state), B);
            e = load_env(v, b);
            hand = hand_to_move(e); last = trick_of(e);
        }
        if (n > 0) {
            Emitter em{actions_u64, off, cap};
            enumerate_legal(masks_of(hand), last, em);
            if (em.pos != off + n && stats) atomicAdd((unsigned long long*)&stats[7], 1ull);  // count/emit disagree
        }
        if (V >= 0 && face) {
            uint64_t pl[C]; float p[2];
            face_planes<(V < 0 ? 0 : V)>(e, pl, p);
#pragma unroll
            for (int c = 0; c < C; c++) s_planes[tid * C + c] = pl[c];
            s_p[tid * 2] = p[0]; s_p[tid * 2 + 1] = p[1];
        }
    }
    __syncthreads();  // packed moves (global) and planes (shared) of the whole CTA are visible

    // ---- face rows: [nenv][C][15] float4, contiguous for the CTA
    if (V >= 0 && face) store_face_rows<C>(face + (size_t)b0 * C * 15, nenv, s_planes, s_p);
    // ---- action rows: [total][15] float4, contiguous for the CTA
    if (actions_f32) {
        long long lim = cap - base; if (lim > total) lim = total; if (lim < 0) lim = 0;
        float4* dst = actions_f32 + (size_t)base * 15;
        const uint64_t* src = actions_u64 + base;
        const long long nvec = lim * 15;
        for (long long i = tid; i < nvec; i += kEnvs) {
            int row = (int)(i / 15), rank = (int)(i - (long long)row * 15);
            dst[i] = thermo_row(__ldcg(src + row), rank, 1.f);
        }
    }
}

__global__ void __launch_bounds__(256) k_encode_actions(const uint64_t* __restrict__ actions, long long n,
                                                        float4* __restrict__ out) {
    long long nvec = n * 15;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        long long row = i / 15; int rank = (int)(i - row * 15);
        out[i] = thermo_row(actions[row], rank, 1.f);
    }
}

}  // namespace ddz

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
using namespace ddz;

static thread_local char g_err[256] = "";
static int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
    return DDZ_E_CUDA;
}
#define DDZ_LAUNCH_CHECK(what)                                         \
    do {                                                               \
        cudaError_t e_ = cudaGetLastError();                           \
        if (e_ != cudaSuccess) return cuda_fail(e_, what);             \
    } while (0)

extern "C" {

int ddz_abi_version(void) { return DDZ_ABI_VERSION; }
int ddz_face_channels(int variant) {
    static const int C[4] = {4, 7, 9, 6};
    return (variant < 0 || variant > 3) ? DDZ_E_ARG : C[variant];
}
size_t ddz_state_bytes(int B) { return B <= 0 ? 0 : (size_t)B * (9 * 8 + 4); }
size_t ddz_workspace_bytes(int B) {
    return B <= 0 ? 0 : align_up((size_t)B * 4, 256) + align_up((size_t)nblocks(B) * 4, 256);
}
const char* ddz_last_error(void) { return g_err; }

int ddz_reset(void* state, const int8_t* perm, const int8_t* lord_pile, int pool_games, int only_done,
              int64_t* stats, int B, void* stream) {
    if (!state || !perm || B <= 0 || pool_games < 1) return DDZ_E_ARG;
    k_reset<<<nblocks(B), kEnvs, 0, (cudaStream_t)stream>>>(state, perm, lord_pile, pool_games, only_done, stats, B);
    DDZ_LAUNCH_CHECK("k_reset");
    return 0;
}

static int launch_emit(const void* state, const uint64_t* hands, const uint64_t* lasts, Workspace ws, int variant,
                       int32_t* offsets, uint64_t* au, float* af, int64_t cap, float* face, int64_t* stats, int B,
                       cudaStream_t st) {
    dim3 g(nblocks(B)), t(kEnvs);
    float4* af4 = (float4*)af; float4* f4 = (float4*)face;
    if (hands) k_emit<-1, true><<<g, t, 0, st>>>(nullptr, hands, lasts, ws, offsets, au, af4, cap, nullptr, stats, B);
    else if (!face) k_emit<-1, false><<<g, t, 0, st>>>(state, nullptr, nullptr, ws, offsets, au, af4, cap, nullptr, stats, B);
    else switch (variant) {
        case 0: k_emit<0, false><<<g, t, 0, st>>>(state, nullptr, nullptr, ws, offsets, au, af4, cap, f4, stats, B); break;
        case 1: k_emit<1, false><<<g, t, 0, st>>>(state, nullptr, nullptr, ws, offsets, au, af4, cap, f4, stats, B); break;
        case 2: k_emit<2, false><<<g, t, 0, st>>>(state, nullptr, nullptr, ws, offsets, au, af4, cap, f4, stats, B); break;
        default: k_emit<3, false><<<g, t, 0, st>>>(state, nullptr, nullptr, ws, offsets, au, af4, cap, f4, stats, B); break;
    }
    DDZ_LAUNCH_CHECK("k_emit");
    return 0;
}

int ddz_observe(const void* state, void* workspace, int variant, int32_t* offsets, uint64_t* actions_u64,
                float* actions_f32, int64_t cap, float* face, int64_t* stats, int B, void* stream) {
    if (!state || !workspace || !offsets || !actions_u64 || B <= 0 || cap < 0) return DDZ_E_ARG;
    if (face && ddz_face_channels(variant) < 0) return DDZ_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    Workspace ws = ws_of(workspace, B);
    StepArgs a; memset(&a, 0, sizeof a);
    k_transition<false, false><<<nblocks(B), kEnvs, 0, st>>>(const_cast<void*>(state), nullptr, nullptr, a, ws, stats, B);
    DDZ_LAUNCH_CHECK("k_transition");
    return launch_emit(state, nullptr, nullptr, ws, variant, offsets, actions_u64, actions_f32, cap, face, stats, B, st);
}

static int fill_step_args(StepArgs& a, const int32_t* offsets, const uint64_t* actions, const void* choice, int mode,
                          uint64_t seed, uint64_t env0, uint32_t stepno, const int32_t rewards[3], int8_t* r,
                          uint8_t* done, int8_t* cat, float* reward) {
    static const int32_t defR[3] = {50, 100, 50};   // game.py:13-14 (up, lord, down)
    if (!offsets || !actions || mode < 0 || mode > 3) return DDZ_E_ARG;
    if (mode != DDZ_CHOICE_PHILOX && !choice) return DDZ_E_ARG;
    memset(&a, 0, sizeof a);
    a.offsets = offsets; a.actions = actions; a.choice = choice; a.mode = mode;
    a.seed = seed; a.env0 = env0; a.stepno = stepno;
    for (int i = 0; i < 3; i++) a.rewards[i] = rewards ? rewards[i] : defR[i];
    a.r = r; a.done = done; a.cat = cat; a.reward = reward;
    a.pool_games = 1;
    return 0;
}

int ddz_step(void* state, const int32_t* offsets, const uint64_t* actions_u64, const void* choice,
             int choice_mode, uint64_t seed, uint64_t env0, uint32_t stepno, const int32_t rewards[3],
             int8_t* r, uint8_t* done, int8_t* cat, float* reward, int64_t* stats, int B, void* stream) {
    if (!state || B <= 0) return DDZ_E_ARG;
    StepArgs a;
    int rc = fill_step_args(a, offsets, actions_u64, choice, choice_mode, seed, env0, stepno, rewards, r, done, cat, reward);
    if (rc) return rc;
    Workspace ws; ws.counts = nullptr; ws.blk = nullptr;
    k_transition<true, false><<<nblocks(B), kEnvs, 0, (cudaStream_t)stream>>>(state, nullptr, nullptr, a, ws, stats, B);
    DDZ_LAUNCH_CHECK("k_transition");
    return 0;
}

int ddz_rollout_step_begin(void* state, void* workspace,
                           const int32_t* prev_offsets, const uint64_t* prev_actions_u64,
                           const void* choice, int choice_mode, uint64_t seed, uint64_t env0, uint32_t stepno,
                           const int32_t rewards[3], const int8_t* perm, const int8_t* lord_pile, int pool_games,
                           int8_t* r, uint8_t* done, int8_t* cat, float* reward, int64_t* stats, int B, void* stream) {
    if (!state || !workspace || B <= 0) return DDZ_E_ARG;
    if (perm && pool_games < 1) return DDZ_E_ARG;
    StepArgs a;
    int rc = fill_step_args(a, prev_offsets, prev_actions_u64, choice, choice_mode, seed, env0, stepno, rewards, r, done, cat, reward);
    if (rc) return rc;
    a.perm = perm; a.lord_pile = lord_pile; a.pool_games = perm ? pool_games : 1;
    k_transition<true, false><<<nblocks(B), kEnvs, 0, (cudaStream_t)stream>>>(state, nullptr, nullptr, a, ws_of(workspace, B), stats, B);
    DDZ_LAUNCH_CHECK("k_transition");
    return 0;
}

int ddz_rollout_step_end(const void* state, void* workspace, int variant,
                         int32_t* out_offsets, uint64_t* out_actions_u64, float* out_actions_f32, int64_t cap,
                         float* face, int64_t* stats, int B, void* stream) {
    if (!state || !workspace || !out_offsets || !out_actions_u64 || B <= 0 || cap < 0) return DDZ_E_ARG;
    if (face && ddz_face_channels(variant) < 0) return DDZ_E_ARG;
    return launch_emit(state, nullptr, nullptr, ws_of(workspace, B), variant, out_offsets, out_actions_u64,
                       out_actions_f32, cap, face, stats, B, (cudaStream_t)stream);
}

int ddz_rollout_step(void* state, void* workspace, int variant,
                     const int32_t* prev_offsets, const uint64_t* prev_actions_u64,
                     const void* choice, int choice_mode, uint64_t seed, uint64_t env0, uint32_t stepno,
                     const int32_t rewards[3], const int8_t* perm, const int8_t* lord_pile, int pool_games,
                     int8_t* r, uint8_t* done, int8_t* cat, float* reward,
                     int32_t* out_offsets, uint64_t* out_actions_u64, float* out_actions_f32, int64_t cap,
                     float* face, int64_t* stats, int B, void* stream) {
    if (!out_offsets || !out_actions_u64 || out_offsets == prev_offsets || out_actions_u64 == prev_actions_u64) return DDZ_E_ARG;
    if (face && ddz_face_channels(variant) < 0) return DDZ_E_ARG;
    int rc = ddz_rollout_step_begin(state, workspace, prev_offsets, prev_actions_u64, choice, choice_mode, seed, env0,
                                    stepno, rewards, perm, lord_pile, pool_games, r, done, cat, reward, stats, B, stream);
    if (rc) return rc;
    return ddz_rollout_step_end(state, workspace, variant, out_offsets, out_actions_u64, out_actions_f32, cap, face,
                                stats, B, stream);
}

int ddz_legal_moves(const uint64_t* hands, const uint64_t* lasts, void* workspace, int32_t* offsets,
                    uint64_t* actions_u64, int64_t cap, int64_t* stats, int n, void* stream) {
    if (!hands || !lasts || !workspace || !offsets || !actions_u64 || n <= 0 || cap < 0) return DDZ_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    Workspace ws = ws_of(workspace, n);
    StepArgs a; memset(&a, 0, sizeof a);
    k_transition<false, true><<<nblocks(n), kEnvs, 0, st>>>(nullptr, hands, lasts, a, ws, stats, n);
    DDZ_LAUNCH_CHECK("k_transition");
    return launch_emit(nullptr, hands, lasts, ws, -1, offsets, actions_u64, nullptr, cap, nullptr, stats, n, st);
}

int ddz_encode_actions(const uint64_t* actions_u64, int64_t n, float* out, void* stream) {
    if (n < 0 || (n > 0 && (!actions_u64 || !out))) return DDZ_E_ARG;
    if (n == 0) return 0;
    long long nvec = (long long)n * 15;
    long long blocks = (nvec + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_encode_actions<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(actions_u64, n, (float4*)out);
    DDZ_LAUNCH_CHECK("k_encode_actions");
    return 0;
}

int ddz_encode_face(const void* state, int variant, float* face, int B, void* stream) {
    if (!state || !face || B <= 0 || ddz_face_channels(variant) < 0) return DDZ_E_ARG;
    dim3 g(nblocks(B)), t(kEnvs);
    cudaStream_t st = (cudaStream_t)stream;
    switch (variant) {
        case 0: k_face<0><<<g, t, 0, st>>>(state, (float4*)face, B); break;
        case 1: k_face<1><<<g, t, 0, st>>>(state, (float4*)face, B); break;
        case 2: k_face<2><<<g, t, 0, st>>>(state, (float4*)face, B); break;
        default: k_face<3><<<g, t, 0, st>>>(state, (float4*)face, B); break;
    }
    DDZ_LAUNCH_CHECK("k_face");
    return 0;
}

}  // extern "C"
