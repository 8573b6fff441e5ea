// ev2b.cu -- host side of libev2b.so: handle, scenario packing, kernel launches (C ABI of include/ev2b.h).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
// (-fmad=false: the battery update must round like the reference's float64 Python arithmetic.)
#include "ev2b_device.cuh"
#include "ev2b_evlist.cuh"
#include "ev2b_spawn.cuh"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

using namespace ev2b;

namespace {

thread_local std::string g_create_error;

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc(&p, count * sizeof(T));
    }
    cudaError_t upload(const std::vector<T> &v) {
        cudaError_t e = alloc(v.size());
        if (e != cudaSuccess || v.empty()) return e;
        return cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { release(); }
};

}  // namespace

struct ev2b_handle {
    ev2b_dims dims{};
    int device = 0;
    int C = 0, P = 0, Tr = 0, T = 0, E = 0, D = 0, EPB = 1, block = 32, n_cls = 1, np_uniform = 0, cs_uniform = 0;
    size_t smem = 0;
    std::string err;
    int64_t launches = 0;
    int64_t launches_by[3] = {0, 0, 0};   // step_kernel, evl_step_kernel, evl_rebuild_kernel (ev2b_kernel_launches)
    const float *last_obs = nullptr;    // obs buffer whose rows are known to be current (incremental obs writes)
    const uint8_t *last_mask = nullptr; // action_mask buffer whose rows are known to be current (evl_step_kernel)
    // host copies of the static layout (needed to pack scenarios)
    std::vector<CsStatic> cs_h;
    std::vector<int> port_cs;           // port -> charger
    std::vector<double> cls_imax;       // per charger class
    std::vector<std::array<double, 4>> cls_veff;
    // device: static
    DevBuf<CsStatic> cs; DevBuf<int> cs_tr_d; DevBuf<int> tr_cs_off, tr_cs_idx, obs_slot, tr_obs_off, port_cs_d, series_off;
    int W = 0;                          // (scenario, time)-only observation values per env
    int series_pairs = 0;               // they can be copied two at a time (Params::series_pairs)
    int obs_pairs = 0;                  // an EV's observation tuple is two floats at an even offset (Params::obs_pairs)
    // device: bank
    int S = 0, Smax = 1, n_dr = 1, lut_len = 101;
    DevBuf<EnvT> env_t; DevBuf<TrT> tr_t; DevBuf<SessRec> sess; DevBuf<EvSpec> spec;
    DevBuf<double> luts_c, luts_d, pot_kw; DevBuf<float> trA, trF, tr_limit, obs_static; DevBuf<DrEv> dr; DevBuf<uint8_t> dr_count;
    // device: state
    DevBuf<uint4> hot; DevBuf<double> cap, exch; DevBuf<int> env_step, env_scn;
    DevBuf<double> env_pot, env_usage, env_kpi, env_pot_prev;
    // device: statistics mode
    int L = 1;
    DevBuf<double> st_soc_sum, st_abs_e, st_act, st_r, cs_sat_sum, cs_dcal, cs_dcyc;
    DevBuf<int> st_cnt, st_nfin, cs_served, cs_em;
    // device: distribution grid
    int n_bus = 0; double s_base = 1000.0;
    DevBuf<double2> grid_Kt, grid_L; DevBuf<double> grid_act, grid_rea, date_feat;
    // stock heuristic agents (ev2b_agent_actions): RoundRobin queue state, action scratch of ev2b_step_k
    double rr_avg_power = 1.0, rr_share = 1.0;
    DevBuf<int> rr_key, rr_fb; DevBuf<double> agent_act;
    // e2e staging (ev2b_step_host)
    DevBuf<unsigned char> st_actions; DevBuf<double> st_reward; DevBuf<uint32_t> st_status; DevBuf<float> st_obs;
    DevBuf<int> st_scn;
    static constexpr int kChunks = 8;   // ev2b_step_host pipelines H2D / kernel / D2H over up to kChunks env chunks
    int n_chunks = 2;                   // chunks in use (measured: 2 -> 285 us, 4 -> 297 us, 8 -> 342 us per c3 step;
                                        // EV2B_HOST_CHUNKS overrides, tuning only)
    cudaStream_t chunk_stream[kChunks] = {};
    cudaEvent_t chunk_ev[kChunks] = {}, start_ev = nullptr;
    // event-driven step kernel (ev2b_evlist.cuh); chosen per handle by EV2B_KERNEL=evlist, used for the launches it covers
    bool evl = false;                   // lists allocated, schedule built
    bool list_valid = true;             // occ_list / occ_n agree with the hot words of every env
    int evl_G = 4, evl_o[14] = {0};     // warps per env; smem map (v_stride, v_amp, ...)
    int evl_tpb = kEvlThreads;          // threads per CTA: 128, or 32 (one env per CTA) for launches of more than one wave
    int n_sm_create = 148;              // SMs of the device (launch-shape decisions at create time)
    bool evl_mix = false;               // EV2B_EVL_MIX=1 (tests): launches that ask for port_energy take step_kernel
    size_t evl_smem = 0;
    std::set<const void *> smem_opted;  // kernels whose dynamic shared-memory limit has been raised on this device
    DevBuf<uint16_t> occ_list; DevBuf<int> occ_n; DevBuf<unsigned> arr_list;
    // device-side scenario sampling (ev2b_spawn.cuh): tables, scratch, the EV-model spec table that replaces the bank's
    bool spawn_ready = false, spawn_specs_live = false, spawn_setpoints = false;
    int spawn_smax = 0, spawn_M = 0;
    SpawnParams spawn_p{};
    DevBuf<double> sp_arr_week, sp_arr_weekend, sp_req, sp_stay, sp_cdf, sp_B, sp_luts, sp_pot_kw;
    DevBuf<int> sp_lut, sp_start, sp_raw_n, sp_n_sess;
    DevBuf<SessRec> sp_raw; DevBuf<EvSpec> sp_spec;
    bool evl_heavy_layout() const { return (dims.flags & EV2B_F_STATS) != 0 || n_bus > 0; }
    void layout_evl() {
        size_t off = 0;
        auto take = [&](size_t bytes, size_t align) { off = (off + align - 1) / align * align; const size_t at = off; off += bytes; return (int)at; };
        const size_t pp = (cs_uniform && np_uniform == 1) ? 0 : (size_t)P;   // one port per charger: no per-port staging
        take(8 * pp, 16);                                          // pw
        evl_o[1] = take(8 * pp, 8); evl_o[2] = take(8 * pp, 8); evl_o[3] = take(8 * (size_t)C, 8);   // amp, pot, csP
        evl_o[4] = take(8 * (size_t)(kEvlTr + 4 * Tr), 16);        // pre (cp.async 16 B destinations)
        evl_o[5] = take(8 * (size_t)EvlNSum * evl_G, 8);           // wsum
        evl_o[7] = 0;                                              // (stage: gone, the list is rebuilt from the occ flags)
        evl_o[8] = take(((size_t)P + 3) / 4 * 4, 4);               // occ
        if (n_bus > 0) { evl_o[6] = take(8 * (size_t)Tr, 8); evl_o[9] = take(16 * 3 * (size_t)n_bus, 16); }   // trp, pfv
        if ((dims.flags & EV2B_F_STATS) && pp) {                   // dsat, dcal, dcyc
            evl_o[10] = take(8 * pp, 8); evl_o[11] = take(8 * pp, 8); evl_o[12] = take(8 * pp, 8);
        }
        evl_o[13] = take(16, 16);                                  // hdr
        evl_o[0] = (int)((off + 15) / 16 * 16);                    // stride
        evl_smem = (size_t)evl_o[0] * (evl_tpb / (32 * evl_G));
    }
    // shared-memory map of step_kernel: byte offsets handed to the kernel through Params (constant bank)
    int so[15] = {0}, pre_stride = 0;
    void layout_smem() {
        const size_t PP = (size_t)EPB * P, nt = (size_t)block, epb = (size_t)EPB;
        size_t off = 0;
        auto take = [&](size_t bytes, size_t align) { off = (off + align - 1) / align * align; const size_t at = off; off += bytes; return (int)at; };
        pre_stride = kPreTr + 4 * Tr;
        take(8 * kNRed * nt, 16);                                  // red
        so[0] = take(8 * PP, 8); so[1] = take(8 * PP, 8); so[2] = take(8 * PP, 8);          // resE, resA, resC
        so[3] = take(8 * epb * Tr, 8); so[4] = take(8 * epb * Tr, 8); so[5] = take(8 * epb, 8);   // trov, trp, lossv
        so[6] = take(16 * epb * 3 * n_bus, 16);                    // pfv
        so[7] = take(8 * epb * kNRed, 8);                          // envs
        so[14] = take(8 * epb * pre_stride, 16);                   // pre
        so[8] = take(8 * PP, 8);                                   // whot
        so[9] = take(4 * nt, 4); so[10] = take(32 * epb, 4); so[11] = take(4 * PP, 4); so[12] = take(16, 4);   // cnt, envi, wl, wcnt
        so[13] = take(PP, 1);                                      // pflag
        smem = off;
    }
    int fail(int code, const char *fmt, ...) {
        char buf[512];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        err = buf;
        return code;
    }
    Params params() const {
        Params p{};
        p.E = E; p.C = C; p.P = P; p.Tr = Tr; p.T = T; p.D = D; p.EPB = EPB; p.n_dr = n_dr; p.lut_len = lut_len;
        p.Smax = Smax; p.S = S; p.n_cls = n_cls; p.W = W; p.env0 = 0; p.env_end = E; p.series_pairs = series_pairs;
        p.reward_kind = dims.reward_kind; p.state_kind = dims.state_kind; p.dr_steps_ahead = dims.dr_steps_ahead;
        p.c60 = 60.0 / (double)dims.timescale; p.p60 = (double)dims.timescale / 60.0; p.period = (double)dims.timescale;
        p.rc60 = 1.0 / p.c60; p.rp60 = 1.0 / p.p60; p.rperiod = 1.0 / p.period;
        p.p_magic = P > 1 ? (unsigned)(0xFFFFFFFFu / (unsigned)P) + 1u : 0u;
        p.c_magic = C > 1 ? (unsigned)(0xFFFFFFFFu / (unsigned)C) + 1u : 0u;
        p.cs_uniform = cs_uniform; p.cs0 = cs_h.empty() ? CsStatic{} : cs_h[0];
        p.o_resE = so[0]; p.o_resA = so[1]; p.o_resC = so[2]; p.o_trov = so[3]; p.o_trp = so[4]; p.o_lossv = so[5];
        p.o_pfv = so[6]; p.o_envs = so[7]; p.o_whot = so[8]; p.o_cnt = so[9]; p.o_envi = so[10]; p.o_wl = so[11];
        p.o_wcnt = so[12]; p.o_pflag = so[13]; p.o_pre = so[14]; p.pre_stride = pre_stride;
        p.n_bus = n_bus; p.s_base = s_base; p.grid_Kt = grid_Kt.p; p.grid_L = grid_L.p; p.grid_act = grid_act.p;
        p.grid_rea = grid_rea.p; p.date_feat = date_feat.p;
        p.stats = (dims.flags & EV2B_F_STATS) ? 1 : 0; p.L = L;
        {   // ev.py:456-516 constants
            const double b_age = 2 * 365, d_dist = 15000, Gk = 0.186, b_cap_ah = 2.05, b_cap_kwh = 78;
            const double q_acc = 2 * (b_age * (d_dist / 365) * Gk * b_cap_ah) / b_cap_kwh;
            p.k_cal = 0.75 / std::pow(b_age, 0.25); p.k_exp = std::exp(-6976.0 / 298.15);
            p.k_cyc = 0.5 * b_cap_ah / std::pow(q_acc, 0.5);
        }
        p.st_soc_sum = st_soc_sum.p; p.st_abs_e = st_abs_e.p; p.st_act = st_act.p; p.st_r = st_r.p;
        p.st_cnt = st_cnt.p; p.st_nfin = st_nfin.p; p.cs_sat_sum = cs_sat_sum.p; p.cs_dcal = cs_dcal.p;
        p.cs_dcyc = cs_dcyc.p; p.cs_served = cs_served.p; p.cs_em = cs_em.p;
        p.cs_tr = cs_tr_d.p;
        p.port_cs = port_cs_d.p; p.series_off = series_off.p; p.obs_static = obs_static.p;
        p.cs = cs.p; p.tr_cs_off = tr_cs_off.p; p.tr_cs_idx = tr_cs_idx.p; p.obs_slot = obs_slot.p; p.tr_obs_off = tr_obs_off.p;
        p.env_t = env_t.p; p.tr_t = tr_t.p; p.sess = sess.p; p.spec = spec.p; p.luts_c = luts_c.p; p.luts_d = luts_d.p;
        p.pot_kw = pot_kw.p; p.trA = trA.p; p.trF = trF.p; p.tr_limit = tr_limit.p; p.dr = dr.p; p.dr_count = dr_count.p;
        p.hot = hot.p; p.cap = cap.p; p.exch = exch.p; p.env_step = env_step.p; p.env_scn = env_scn.p;
        p.env_pot = env_pot.p; p.env_usage = env_usage.p; p.env_kpi = env_kpi.p; p.env_pot_prev = env_pot_prev.p;
        p.occ_list = occ_list.p; p.occ_n = occ_n.p; p.arr_list = arr_list.p;
        p.v_stride = evl_o[0]; p.v_amp = evl_o[1]; p.v_pot = evl_o[2]; p.v_csP = evl_o[3]; p.v_pre = evl_o[4];
        { int lg = 0; while ((2 << lg) * Tr <= 32) ++lg; p.tr_lg = lg; }
        p.v_wsum = evl_o[5]; p.v_trp = evl_o[6]; p.v_stage = evl_o[7]; p.v_occ = evl_o[8]; p.v_pfv = evl_o[9];
        p.v_dsat = evl_o[10]; p.v_dcal = evl_o[11]; p.v_dcyc = evl_o[12]; p.v_hdr = evl_o[13];
        {   // auto-reset stride: E mod S (so that with E < S consecutive episodes walk the bank), made coprime to S --
            // when E is a multiple of S that is 1: every env visits every scenario instead of replaying its first one
            int st = S > 0 ? E % S : 0;
            if (st == 0) st = 1;
            while (S > 1 && std::gcd(st, S) != 1) ++st;
            p.scn_stride = st;
        }
        p.rr_key = rr_key.p; p.rr_fb = rr_fb.p; p.rr_avg_power = rr_avg_power; p.rr_share = rr_share;
        return p;
    }
};

#define CUDA_TRY(h, expr)                                                                       \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) return (h)->fail(EV2B_E_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

static int obs_dim_for(int kind, int P, int Tr) {
    switch (kind) {
    case EV2B_STATE_PUBLIC_PST: return 3 + 3 * P;
    case EV2B_STATE_V2G_PROFIT_MAX: return 22 + 2 * P;
    case EV2B_STATE_V2G_PROFIT_MAX_LOADS: return 22 + 40 * Tr + 2 * P;
    case EV2B_STATE_V2G_GRID: return 6 + 2 * Tr + 3 * P;          // n_bus == Tr in grid mode
    default: return 0;
    }
}

// The full-featured (HEAVY) instantiation: statistics mode, distribution grid, the less common rewards.
static bool needs_heavy(const ev2b_handle *h) {
    return (h->dims.flags & EV2B_F_STATS) != 0 || h->n_bus > 0 || h->dims.reward_kind >= EV2B_REWARD_SQTR_TR_USER;
}
// Does the event-driven kernel serve this handle's launches?  (it covers every feature; the choice is made at create time)
static bool evl_covers(const ev2b_handle *h, const ev2b_step_out *o) {
    if (h->evl_mix && o && o->port_energy) return false;   // test-only (EV2B_EVL_MIX=1): exercise both kernels on one handle
    return h->evl;
}

// Raises a kernel's dynamic shared-memory limit once per handle (not on every launch).
template <typename K>
static cudaError_t opt_in_smem(ev2b_handle *h, K kern, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    const void *key = reinterpret_cast<const void *>(kern);
    if (h->smem_opted.count(key)) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) h->smem_opted.insert(key);
    return e;
}

// kstep: the KSTEP instantiation (p.k_steps steps in one launch; lean kernels only, see ev2b_step_k)
template <typename ActT>
static cudaError_t launch_evl(ev2b_handle *h, const Params &p_in, cudaStream_t st, bool kstep = false) {
    Params p = p_in;
    // two-at-a-time accesses need the caller's buffers on an 8 / 16-byte boundary (cudaMalloc and torch give 256)
    const bool obs_al = (reinterpret_cast<uintptr_t>(p.out.obs) & 7u) == 0;
    p.series_pairs = (h->series_pairs && obs_al) ? 1 : 0;
    p.obs_pairs = (h->obs_pairs && obs_al) ? 1 : 0;
    p.act_pairs = (p.agent_kind == EV2B_AGENT_EXTERNAL && h->cs_uniform && h->np_uniform == 2 && p.actions &&
                   (reinterpret_cast<uintptr_t>(p.actions) & (2 * sizeof(ActT) - 1)) == 0) ? 1 : 0;
    const int epb = h->evl_tpb / (32 * h->evl_G);
    const unsigned grid = (unsigned)((p.env_end - p.env0 + epb - 1) / epb);
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = opt_in_smem(h, kern, h->evl_smem);
        if (e != cudaSuccess) return e;
        EV2B_LAUNCH(kern, grid, h->evl_tpb, h->evl_smem, st, p);
        return cudaGetLastError();
    };
    const int np = (h->cs_uniform && (h->np_uniform == 1 || h->np_uniform == 2)) ? h->np_uniform : 0;
    // HEAVY: statistics mode, the distribution grid, the dense per-port outputs
    const bool heavy = h->evl_heavy_layout() || p.out.dep_sat || p.out.dep_cap || p.out.port_energy || p.out.node_voltage;
#define EV2B_EVL_DISPATCH(G, TPB)                                                          \
    do {                                                                                   \
        if (heavy) {                                                                       \
            if (kstep) return cudaErrorInvalidValue;                                       \
            if (np == 1) return go(evl_step_kernel<ActT, 1, true, G, true, false, TPB>);   \
            if (np == 2) return go(evl_step_kernel<ActT, 2, true, G, true, false, TPB>);   \
            return go(evl_step_kernel<ActT, 0, false, G, true, false, TPB>);               \
        }                                                                                  \
        if (kstep) {                                                                       \
            if (np == 1) return go(evl_step_kernel<ActT, 1, true, G, false, true, TPB>);   \
            if (np == 2) return go(evl_step_kernel<ActT, 2, true, G, false, true, TPB>);   \
            return go(evl_step_kernel<ActT, 0, false, G, false, true, TPB>);               \
        }                                                                                  \
        if (np == 1) return go(evl_step_kernel<ActT, 1, true, G, false, false, TPB>);      \
        if (np == 2) return go(evl_step_kernel<ActT, 2, true, G, false, false, TPB>);      \
        return go(evl_step_kernel<ActT, 0, false, G, false, false, TPB>);                  \
    } while (0)
    if (h->evl_G == 1 && h->evl_tpb == 32) EV2B_EVL_DISPATCH(1, 32);
    if (h->evl_G == 1) EV2B_EVL_DISPATCH(1, kEvlThreads);
    if (h->evl_G == 2) EV2B_EVL_DISPATCH(2, kEvlThreads);
    EV2B_EVL_DISPATCH(4, kEvlThreads);
#undef EV2B_EVL_DISPATCH
}

// Brings occ_list / occ_n back in line with the hot words after step_kernel advanced some envs (all envs, on `st`).
static int ensure_list(ev2b_handle *h, cudaStream_t st) {
    if (!h->evl || h->list_valid) return EV2B_OK;
    const int wpb = 4;
    EV2B_LAUNCH(evl_rebuild_kernel, (unsigned)((h->E + wpb - 1) / wpb), 32 * wpb, 0, st, h->params(), 0, h->E);
    h->launches += 1; h->launches_by[2] += 1;
    if (cudaGetLastError() != cudaSuccess) return h->fail(EV2B_E_CUDA, "evl_rebuild_kernel launch failed");
    h->list_valid = true;
    return EV2B_OK;
}

template <typename ActT>
static cudaError_t launch_step(ev2b_handle *h, const Params &p, cudaStream_t st) {
    const unsigned grid = (unsigned)((p.env_end - p.env0 + h->EPB - 1) / h->EPB);
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = opt_in_smem(h, kern, h->smem);
        if (e != cudaSuccess) return e;
        EV2B_LAUNCH(kern, grid, h->block, h->smem, st, p);
        return cudaGetLastError();
    };
#ifndef EV2B_MINB
#define EV2B_MINB 4
#endif
    // uniform charger layout + 1 or 2 ports: the YAML case; everything else takes the generic variant
    const int np = (h->cs_uniform && (h->np_uniform == 1 || h->np_uniform == 2)) ? h->np_uniform : 0;
#define EV2B_DISPATCH(MAXT, MINB)                                                       \
    do {                                                                                \
        if (stats) {                                                                    \
            if (np == 1) return go(step_kernel<ActT, 1, true, MAXT, MINB, true, true>);  \
            if (np == 2) return go(step_kernel<ActT, 2, true, MAXT, MINB, true, true>);  \
            return go(step_kernel<ActT, 0, false, MAXT, MINB, true, true>);              \
        }                                                                               \
        if (opt) {                                                                      \
            if (np == 1) return go(step_kernel<ActT, 1, true, MAXT, MINB, false, true>); \
            if (np == 2) return go(step_kernel<ActT, 2, true, MAXT, MINB, false, true>); \
            return go(step_kernel<ActT, 0, false, MAXT, MINB, false, true>);             \
        }                                                                               \
        if (np == 1) return go(step_kernel<ActT, 1, true, MAXT, MINB, false, false>);   \
        if (np == 2) return go(step_kernel<ActT, 2, true, MAXT, MINB, false, false>);   \
        return go(step_kernel<ActT, 0, false, MAXT, MINB, false, false>);               \
    } while (0)
    const bool stats = needs_heavy(h);   // the HEAVY instantiation
#ifndef EV2B_MINB128
#define EV2B_MINB128 8
#endif
    const ev2b_step_out &o = p.out;    // any optional output requested?  (reward / status / obs are always compiled in)
    const bool opt = o.cs_power || o.cs_current || o.tr_power || o.tr_overload || o.total_costs || o.action_mask ||
                     o.dep_sat || o.dep_cap || o.port_energy || o.node_voltage || o.hist_cs_power || o.hist_cs_current ||
                     o.hist_tr_overload || o.hist_usage;
    if (h->block <= 128) EV2B_DISPATCH(128, EV2B_MINB128);
    if (h->block <= 256) EV2B_DISPATCH(256, EV2B_MINB);
    if (h->block <= 512) EV2B_DISPATCH(512, 2);
    EV2B_DISPATCH(kMaxThreads, 1);
#undef EV2B_DISPATCH
}

extern "C" {

int ev2b_abi_version(void) { return EV2B_ABI_VERSION; }

const char *ev2b_last_error(const ev2b_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int ev2b_create(const ev2b_dims *d, const ev2b_topology *tp, int device, ev2b_handle **out) {
    if (!d || !tp || !out) { g_create_error = "null argument"; return EV2B_E_ARG; }
    *out = nullptr;
    if (d->n_envs < 1 || d->n_chargers < 1 || d->n_transformers < 1 || d->sim_length < 1 || d->timescale < 1) {
        g_create_error = "ev2b_create: sizes must be >= 1"; return EV2B_E_ARG;
    }
    if (d->sim_length > 32000) { g_create_error = "ev2b_create: sim_length > 32000 (int16 step fields)"; return EV2B_E_LIMIT; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        g_create_error = "ev2b_create: no CUDA device (this library has no CPU fallback)"; return EV2B_E_CUDA;
    }
    if (cudaSetDevice(device) != cudaSuccess) { g_create_error = "ev2b_create: cudaSetDevice failed"; return EV2B_E_CUDA; }
    ev2b_handle *h = new ev2b_handle();
    h->dims = *d; h->device = device;
    h->E = d->n_envs; h->C = d->n_chargers; h->Tr = d->n_transformers; h->T = d->sim_length;
    const int C = h->C;
    // ---- static charger table ------------------------------------------------------------------
    h->cs_h.resize(C);
    int off = 0;
    bool uniform_ports = true;
    std::map<std::array<double, 3>, int> cls_of;
    double rr_acc = 0.0;
    for (int c = 0; c < C; ++c) {
        CsStatic &s = h->cs_h[c];
        const int ph = tp->cs_phases[c];
        if (ph < 1 || ph > 3 || tp->cs_n_ports[c] < 1 || tp->cs_n_ports[c] > 1023 || tp->cs_tr[c] < 0 ||
            tp->cs_tr[c] >= h->Tr) {
            delete h; g_create_error = "ev2b_create: bad charger entry (phases 1..3, n_ports 1..1023, tr in range)";
            return EV2B_E_ARG;
        }
        s.imax = tp->cs_imax[c]; s.imin = tp->cs_imin[c];
        s.imax_dis_abs = std::fabs(tp->cs_imax_dis[c]);          // ev_charger.py:184
        s.imin_dis = tp->cs_imin_dis[c];
        s.veff[0] = 0.0;
        s.rveff[0] = 0.0;
        for (int k = 1; k <= 3; ++k) {
            s.veff[k] = tp->cs_voltage[c] * std::sqrt((double)k);   // ev.py:279
            s.rveff[k] = 1.0 / s.veff[k];
        }
        s.max_power = std::sqrt((double)ph) * tp->cs_voltage[c] * s.imax / 1000.0;           // utils.py:779-780
        s.min_power = std::sqrt((double)ph) * tp->cs_voltage[c] * s.imin / 1000.0;           // utils.py:781-782
        s.calap_kw = s.imax * tp->cs_voltage[c] * std::sqrt((double)ph) / 1000.0;              // heuristics.py:120-121
        rr_acc += s.imax * tp->cs_voltage[c] * std::sqrt((double)ph) / (double)tp->cs_n_ports[c];   // heuristics.py:20-23
        s.port_off = off; s.n_ports = tp->cs_n_ports[c]; s.tr = tp->cs_tr[c]; s.phases = ph;
        std::array<double, 3> key{s.imax, tp->cs_voltage[c], (double)ph};
        auto it = cls_of.find(key);
        if (it == cls_of.end()) {
            it = cls_of.emplace(key, (int)cls_of.size()).first;
            h->cls_imax.push_back(s.imax);
            h->cls_veff.push_back({s.veff[0], s.veff[1], s.veff[2], s.veff[3]});
        }
        s.cls = it->second;
        off += s.n_ports;
        if (s.n_ports != h->cs_h[0].n_ports) uniform_ports = false;
        for (int j = 0; j < s.n_ports; ++j) h->port_cs.push_back(c);
    }
    h->P = off;
    h->rr_avg_power = rr_acc / (double)C;
    h->rr_share = 1.0 / (double)tp->cs_n_ports[0];                 // 1 / env.number_of_ports_per_cs  heuristics.py:84
    if (tp->n_bus > 0) {
        if (tp->n_bus != h->Tr || tp->n_bus > 128 || !tp->grid_K || !tp->grid_L) {
            delete h; g_create_error = "ev2b_create: grid needs n_bus == n_transformers <= 128 and K, L"; return EV2B_E_ARG;
        }
        h->n_bus = tp->n_bus; h->s_base = tp->grid_s_base;
    }
    if (d->reward_kind < EV2B_REWARD_NONE || d->reward_kind > EV2B_REWARD_SQ_TRACKING_PENALTY) {
        delete h; g_create_error = "ev2b_create: unknown reward kind"; return EV2B_E_ARG;
    }
    if ((d->reward_kind == EV2B_REWARD_GRID_FULL || d->reward_kind == EV2B_REWARD_GRID_SIMPLE ||
         d->reward_kind == EV2B_REWARD_GRID_PROFITMAX_V2 ||
         d->state_kind == EV2B_STATE_V2G_GRID) && h->n_bus == 0) {
        delete h; g_create_error = "ev2b_create: grid reward/state functions need a grid (n_bus > 0)"; return EV2B_E_ARG;
    }
    h->n_cls = (int)cls_of.size();
    h->np_uniform = uniform_ports ? h->cs_h[0].n_ports : 0;
    h->cs_uniform = (uniform_ports && h->n_cls == 1) ? 1 : 0;
    for (int c = 1; c < C && h->cs_uniform; ++c) {
        const CsStatic &a0 = h->cs_h[0], &b0 = h->cs_h[c];
        if (a0.imin != b0.imin || a0.imax_dis_abs != b0.imax_dis_abs || a0.imin_dis != b0.imin_dis) h->cs_uniform = 0;
    }
    h->D = obs_dim_for(d->state_kind, h->P, h->Tr);
    // transformer -> chargers CSR (id order)
    std::vector<int> tr_off(h->Tr + 1, 0), tr_idx(C);
    for (int c = 0; c < C; ++c) tr_off[h->cs_h[c].tr + 1]++;
    for (int k = 0; k < h->Tr; ++k) tr_off[k + 1] += tr_off[k];
    { std::vector<int> fill(tr_off.begin(), tr_off.end() - 1);
      for (int c = 0; c < C; ++c) tr_idx[fill[h->cs_h[c].tr]++] = c; }
    // observation slots: transformer-major, then charger id, then port   state.py:37-42, 85-92, 128-141
    std::vector<int> slot(h->P, 0), tr_obs(h->Tr, 0);
    {
        const int kind = d->state_kind;
        const int tuple = kind == EV2B_STATE_PUBLIC_PST ? 3 : 2;
        int o = kind == EV2B_STATE_PUBLIC_PST ? 3 : 22;
        if (kind == EV2B_STATE_V2G_GRID)      // charger-major, 3 values per port  state.py:262-274
            for (int pp = 0; pp < h->P; ++pp) slot[pp] = 6 + 2 * h->n_bus + 3 * pp;
        else
        for (int k = 0; k < h->Tr; ++k) {
            if (kind == EV2B_STATE_V2G_PROFIT_MAX_LOADS) { tr_obs[k] = o; o += 40; }
            for (int i = tr_off[k]; i < tr_off[k + 1]; ++i) {
                const CsStatic &s = h->cs_h[tr_idx[i]];
                for (int j = 0; j < s.n_ports; ++j) { slot[s.port_off + j] = o; o += tuple; }
            }
        }
    }
    // destination offsets of the (scenario, time)-only observation values: [20 prices][Tr x 40]
    std::vector<int> series_off_h;
    if (d->state_kind == EV2B_STATE_V2G_PROFIT_MAX || d->state_kind == EV2B_STATE_V2G_PROFIT_MAX_LOADS) {
        for (int i = 0; i < 20; ++i) series_off_h.push_back(2 + i);
        if (d->state_kind == EV2B_STATE_V2G_PROFIT_MAX_LOADS)
            for (int k = 0; k < h->Tr; ++k) for (int j = 0; j < 40; ++j) series_off_h.push_back(tr_obs[k] + j);
    }
    if (d->state_kind == EV2B_STATE_V2G_GRID) {
        for (int i = 0; i < 5; ++i) series_off_h.push_back(i);
        for (int i = 0; i < 2 * h->n_bus; ++i) series_off_h.push_back(6 + i);
    }
    h->obs_pairs = (h->D % 2 == 0 && d->state_kind != EV2B_STATE_NONE && d->state_kind != EV2B_STATE_PUBLIC_PST &&
                    d->state_kind != EV2B_STATE_V2G_GRID) ? 1 : 0;          // (those two write three values per EV)
    for (int i = 0; i < h->P && h->obs_pairs; ++i) if (slot[i] % 2 != 0) h->obs_pairs = 0;
    h->W = (int)series_off_h.size();
    h->series_pairs = (h->W > 0 && h->W % 2 == 0 && h->D % 2 == 0) ? 1 : 0;
    for (int i = 0; i + 1 < h->W && h->series_pairs; i += 2)
        if (series_off_h[i] % 2 != 0 || series_off_h[i + 1] != series_off_h[i] + 1) h->series_pairs = 0;
    // launch shape: a CTA owns EPB whole envs, one thread per (env, charger)
    {
        int best_epb = 1; double best_u = -1;
        const int cap_thr = C > 256 ? kMaxThreads : (C > 128 ? 256 : 128);   // small CTAs: less barrier skew (measured)
        int n_sm = 148;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
        h->n_sm_create = n_sm;
        for (int epb = 1; epb <= 64 && epb <= h->E; ++epb) {
            const int thr = epb * C;
            if (thr > cap_thr) break;               // (C > 1024: no step_kernel shape; the event-driven kernel serves the handle)
            // packing several small envs into a CTA fills the last warp, but only pays once the grid still has a few
            // CTAs per SM to overlap their serial phases (c2, 1024 x 25: 8.8 us with 1 env per CTA, 10.1 us with 5)
            if (epb > 1 && (h->E + epb - 1) / epb < 4 * n_sm) break;
            const int blk = (thr + 31) / 32 * 32;
            const double u = (double)thr / blk;
            if (u > best_u + 1e-9) { best_u = u; best_epb = epb; }
        }
        if (const char *ov = getenv("EV2B_EPB")) {          // tuning override (benchmarks only)
            const int v = atoi(ov);
            if (v >= 1 && v <= h->E && v * C <= kMaxThreads) best_epb = v;
        }
        h->EPB = best_epb;
        h->block = std::min(kMaxThreads, std::max(32, (best_epb * C + 31) / 32 * 32));
        h->layout_smem();
    }
    // Which step kernel serves this handle (ev2b_evlist.cuh).  The event-driven kernel is the default from 1024 envs up
    // (B200, whole episodes, us per launch, event-driven vs step_kernel: c3 4096 envs 19.9 vs 37.8; c4 8192 envs 35.3 vs
    // 100.2; at 1024 envs, four warps per env: c3 9.0 vs 10.4, c2 8.6 vs 9.1 -- profiles/r2_ab_small_batches.jsonl; round
    // 1's event-driven kernel lost there, 15.7 vs 13.6).  Smaller batches were not measured and keep step_kernel.
    // EV2B_KERNEL=percharger|evlist forces one, EV2B_EVL_G = warps per env (1, 2, 4) overrides the group size (tuning / tests).
    {
        const char *kv = getenv("EV2B_KERNEL");
        const bool force_on = kv && (!strcmp(kv, "evlist") || !strcmp(kv, "evl"));
        const bool force_off = kv && !strcmp(kv, "percharger");
        const bool big = C > kMaxThreads || h->smem > 200 * 1024;     // beyond step_kernel's one-thread-per-charger shape
        const bool want = force_on || big || (!force_off && h->E >= 1024);
        if (want && h->P < 65535) {
            h->evl = true;
            // warps per env: as many as it takes to fill the machine (~28 resident warps per SM of the lean kernel, 16 of the
            // HEAVY one), at most 4.  B200, whole episodes, us per launch, G = 1 / 2 / 4: c3 (4096 envs) 23.3 / 26.6 / 34.0,
            // c4 (8192) 46.0 / 53.0 / 59.4, c5 (2048, HEAVY) 74.8 / 86.9; c3 at 1024 envs 17.5 / 12.3 (profiles/r2_ab_*.json)
            {
                const int cap_warps = h->n_sm_create * (h->evl_heavy_layout() ? 16 : 28);
                h->evl_G = h->E >= cap_warps / 2 ? 1 : (h->E >= cap_warps / 4 ? 2 : 4);
            }
            if (const char *gv = getenv("EV2B_EVL_G")) { const int v = atoi(gv); if (v == 1 || v == 2 || v == 4) h->evl_G = v; }
            // one env per CTA when the launch is more than the machine holds at once (see evl_step_kernel)
            h->evl_tpb = (h->evl_G == 1 && h->E > h->n_sm_create * (h->evl_heavy_layout() ? 16 : 28)) ? 32 : kEvlThreads;
            if (const char *tv = getenv("EV2B_EVL_TPB")) { const int v = atoi(tv); if (v == kEvlThreads || (v == 32 && h->evl_G == 1)) h->evl_tpb = v; }
            if (const char *mv = getenv("EV2B_EVL_MIX")) h->evl_mix = atoi(mv) != 0 && !big;
            h->layout_evl();
            if (h->evl_smem > 200 * 1024 && h->evl_G == 1) { h->evl_tpb = 32; h->layout_evl(); }     // fewer envs per CTA
            while (h->evl_smem > 200 * 1024 && h->evl_G < 4) { h->evl_G *= 2; h->evl_tpb = kEvlThreads; h->layout_evl(); }
            if (h->evl_smem > 200 * 1024) h->evl = false;          // does not fit: step_kernel takes every launch
        }
        if (!h->evl && big) {
            delete h; g_create_error = "ev2b_create: env too large (per-env working set exceeds shared memory, or >= 65535 ports)";
            return EV2B_E_LIMIT;
        }
    }
#define CREATE_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { \
        g_create_error = std::string(#expr) + ": " + cudaGetErrorString(_e); delete h; return EV2B_E_CUDA; } } while (0)
    CREATE_TRY(h->cs.upload(h->cs_h));
    CREATE_TRY(h->tr_cs_off.upload(tr_off));
    CREATE_TRY(h->tr_cs_idx.upload(tr_idx));
    CREATE_TRY(h->obs_slot.upload(slot));
    CREATE_TRY(h->tr_obs_off.upload(tr_obs));
    CREATE_TRY(h->port_cs_d.upload(h->port_cs));
    { std::vector<int> ct(C); for (int c = 0; c < C; ++c) ct[c] = h->cs_h[c].tr; CREATE_TRY(h->cs_tr_d.upload(ct)); }
    CREATE_TRY(h->series_off.upload(series_off_h));
    if (h->n_bus > 0) {
        const int n = h->n_bus;
        std::vector<double2> kt((size_t)n * n), lv(n);
        for (int r = 0; r < n; ++r) {
            for (int c2 = 0; c2 < n; ++c2) kt[(size_t)c2 * n + r] = make_double2(tp->grid_K[2 * ((size_t)r * n + c2)], tp->grid_K[2 * ((size_t)r * n + c2) + 1]);
            lv[r] = make_double2(tp->grid_L[2 * r], tp->grid_L[2 * r + 1]);
        }
        CREATE_TRY(h->grid_Kt.upload(kt)); CREATE_TRY(h->grid_L.upload(lv));
    }
    const size_t EP = (size_t)h->E * h->P;
    CREATE_TRY(h->hot.alloc(EP)); CREATE_TRY(h->cap.alloc(EP)); CREATE_TRY(h->exch.alloc(EP));
    CREATE_TRY(h->env_step.alloc(h->E)); CREATE_TRY(h->env_scn.alloc(h->E));
    CREATE_TRY(h->env_pot.alloc(h->E)); CREATE_TRY(h->env_usage.alloc(h->E)); CREATE_TRY(h->env_pot_prev.alloc(h->E));
    CREATE_TRY(h->env_kpi.alloc((size_t)h->E * EV2B_KPI_COUNT));
    if (h->evl) {
        CREATE_TRY(h->occ_list.alloc(EP)); CREATE_TRY(h->occ_n.alloc(h->E));
        CREATE_TRY(cudaMemset(h->occ_list.p, 0, EP * sizeof(uint16_t)));
        CREATE_TRY(cudaMemset(h->occ_n.p, 0, h->E * sizeof(int)));
    }
    CREATE_TRY(cudaMemset(h->hot.p, 0, EP * sizeof(uint4)));
    CREATE_TRY(cudaMemset(h->cap.p, 0, EP * sizeof(double)));
    CREATE_TRY(cudaMemset(h->exch.p, 0, EP * sizeof(double)));
    CREATE_TRY(cudaMemset(h->env_scn.p, 0, h->E * sizeof(int)));
    CREATE_TRY(cudaMemset(h->env_pot.p, 0, h->E * sizeof(double)));
    CREATE_TRY(cudaMemset(h->env_usage.p, 0, h->E * sizeof(double)));
    CREATE_TRY(cudaMemset(h->env_kpi.p, 0, (size_t)h->E * EV2B_KPI_COUNT * sizeof(double)));
    {   // until reset() is called every env reads as "done"
        std::vector<int> st(h->E, h->T);
        CREATE_TRY(cudaMemcpy(h->env_step.p, st.data(), st.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
#undef CREATE_TRY
    *out = h;
    return EV2B_OK;
}

void ev2b_destroy(ev2b_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (int i = 0; i < ev2b_handle::kChunks; ++i) {
        if (h->chunk_stream[i]) cudaStreamDestroy(h->chunk_stream[i]);
        if (h->chunk_ev[i]) cudaEventDestroy(h->chunk_ev[i]);
    }
    if (h->start_ev) cudaEventDestroy(h->start_ev);
    delete h;
}

int ev2b_obs_dim(const ev2b_handle *h) { return h ? h->D : 0; }
int ev2b_n_ports(const ev2b_handle *h) { return h ? h->P : 0; }
int ev2b_n_scenarios(const ev2b_handle *h) { return h ? h->S : 0; }
int64_t ev2b_launch_count(const ev2b_handle *h) { return h ? h->launches : 0; }
int64_t ev2b_kernel_launches(const ev2b_handle *h, int which) { return (h && which >= 0 && which < 3) ? h->launches_by[which] : 0; }

// k/1000.0 == v  <=>  v is what np.round(x, 3) produces (utils.py:293-296, 309-310)
static bool milli(double v, unsigned *k) {
    const double r = std::rint(v * 1000.0);
    if (!(r >= 0.0 && r < 65535.0)) return false;
    if (r / 1000.0 != v) return false;
    *k = (unsigned)r;
    return true;
}

int ev2b_load_scenarios(ev2b_handle *h, const ev2b_scenarios *b) {
    if (!h || !b) return EV2B_E_ARG;
    if (b->n < 1) return h->fail(EV2B_E_ARG, "load_scenarios: empty bank");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int S = b->n, P = h->P, T = h->T, Tr = h->Tr, C = h->C;
    const int n_dr = std::max(1, b->n_dr);
    const int lut_len = b->lut_len > 0 ? b->lut_len : 101;

    // ---- efficiency tables: global de-duplication ------------------------------------------------
    std::vector<double> luts_c, luts_d;
    std::map<std::string, int> lut_ids;
    std::vector<int> lut_global((size_t)b->lut_off[S], -1);
    for (int64_t i = 0; i < b->lut_off[S]; ++i) {
        std::string key((const char *)(b->luts_c + i * lut_len), sizeof(double) * lut_len);
        key.append((const char *)(b->luts_d + i * lut_len), sizeof(double) * lut_len);
        auto it = lut_ids.find(key);
        if (it == lut_ids.end()) {
            it = lut_ids.emplace(key, (int)lut_ids.size()).first;
            luts_c.insert(luts_c.end(), b->luts_c + i * lut_len, b->luts_c + (i + 1) * lut_len);
            luts_d.insert(luts_d.end(), b->luts_d + i * lut_len, b->luts_d + (i + 1) * lut_len);
        }
        lut_global[i] = it->second;
    }

    // ---- sessions: replay evs_connected.index(None) (ev_charger.py:273) and pack ----------------
    std::vector<EvSpec> specs;
    std::map<std::string, int> spec_ids;
    struct Placed { int port; int64_t row; };
    std::vector<std::vector<Placed>> placed(S);
    int Smax = 1;
    std::vector<int> occupied_until(P);
    std::vector<int> per_port(P);
    for (int i = 0; i < S; ++i) {
        std::fill(occupied_until.begin(), occupied_until.end(), -1);
        std::fill(per_port.begin(), per_port.end(), 0);
        int prev_arr = 1;
        for (int64_t r = b->sess_off[i]; r < b->sess_off[i + 1]; ++r) {
            const int ta = b->s_t_arr[r], td = b->s_t_dep[r], loc = b->s_loc[r];
            if (ta < prev_arr) return h->fail(EV2B_E_SCENARIO, "scenario %d: sessions must be arrival-sorted with t_arr >= 1", i);
            prev_arr = ta;
            if (loc < 0 || loc >= C) return h->fail(EV2B_E_SCENARIO, "scenario %d: session location %d out of range", i, loc);
            if (td < ta) return h->fail(EV2B_E_SCENARIO, "scenario %d: departure before arrival", i);
            const CsStatic &cs = h->cs_h[loc];
            int port = -1;
            for (int j = 0; j < cs.n_ports; ++j) {
                int &ou = occupied_until[cs.port_off + j];
                if (ou >= 0 && ou <= ta - 1) ou = -1;     // left during a step <= ta-1
                if (ou < 0 && port < 0) port = cs.port_off + j;
            }
            if (port < 0) return h->fail(EV2B_E_SCENARIO, "scenario %d: charger %d has no free port at step %d", i, loc, ta);
            occupied_until[port] = td;
            placed[i].push_back({port, r});
            Smax = std::max(Smax, ++per_port[port]);
        }
    }
    if (h->spawn_ready) Smax = std::max(Smax, h->spawn_smax);       // room for whatever the device sampler can draw
    if (Smax > 255) return h->fail(EV2B_E_LIMIT, "more than 255 sessions on one port");
    int Lmax = 2;
    SessRec empty{};
    empty.hot.x = ((unsigned)kNoArrival & 0xFFFFu) | (0xFFFFu << 16);
    empty.hot.y = (unsigned)kNoArrival & 0xFFFFu;
    std::vector<SessRec> sess((size_t)S * P * Smax, empty);
    std::vector<int> last_of_port(P);
    // arrival schedule of the event-driven kernel: sessions of scenario i arriving at step q (q <= T), in session order
    std::vector<int> arr_off_h((size_t)S * (T + 2), 0);
    std::vector<unsigned> arr_list_h;
    for (int i = 0; i < S; ++i) {
        std::fill(per_port.begin(), per_port.end(), 0);
        std::fill(last_of_port.begin(), last_of_port.end(), -1);
        const size_t arr_base = arr_list_h.size();
        for (const Placed &pl : placed[i]) {
            const int64_t r = pl.row;
            EvSpec sp{};
            sp.B = b->s_B[r]; sp.pmax_ac = b->s_pmax_ac[r]; sp.pmin_ac = b->s_pmin_ac[r];
            sp.pmax_dis = b->s_pmax_dis[r]; sp.pmin_dis = b->s_pmin_dis[r]; sp.bmin = b->s_bmin[r];
            sp.bmin_em = b->s_bmin_em[r]; sp.desired = b->s_desired[r]; sp.mult = b->s_mult[r];
            sp.ev_phases = b->s_ev_phases[r];
            sp.rB = 1.0 / sp.B;
            if (sp.ev_phases < 1 || sp.ev_phases > 3) return h->fail(EV2B_E_SCENARIO, "scenario %d: ev_phases must be 1..3", i);
            sp.lut = b->s_lut[r] >= 0 ? lut_global[b->lut_off[i] + b->s_lut[r]] : -1;
            unsigned tsm = 0xFFFFu, ecm = 0xFFFFu, edm = 0xFFFFu;
            sp.ts = sp.eta_c = sp.eta_d = 0.0;
            if (!milli(b->s_ts[r], &tsm)) { tsm = 0xFFFFu; sp.ts = b->s_ts[r]; }
            if (sp.lut < 0) {
                if (!milli(b->s_eta_c[r], &ecm)) { ecm = 0xFFFFu; sp.eta_c = b->s_eta_c[r]; }
                if (!milli(b->s_eta_d[r], &edm)) { edm = 0xFFFFu; sp.eta_d = b->s_eta_d[r]; }
                if (!(b->s_eta_c[r] > 0.0)) return h->fail(EV2B_E_SCENARIO, "scenario %d: charge_efficiency must be > 0 (ev.py:293)", i);
            } else { ecm = edm = 0; }
            std::string key((const char *)&sp, sizeof sp);
            auto it = spec_ids.find(key);
            if (it == spec_ids.end()) {
                if (specs.size() >= 65535) return h->fail(EV2B_E_LIMIT, "more than 65535 distinct EV specs");
                it = spec_ids.emplace(key, (int)specs.size()).first;
                specs.push_back(sp);
            }
            const int ta = std::min(b->s_t_arr[r], 32766), td = std::min(b->s_t_dep[r], 32766);
            const int k = per_port[pl.port]++;
            SessRec &rec = sess[((size_t)i * P + pl.port) * Smax + k];
            rec.hot.x = ((unsigned)ta & 0xFFFFu) | (((unsigned)td & 0xFFFFu) << 16);
            rec.hot.y = ((unsigned)kNoArrival & 0xFFFFu) | ((unsigned)(k + 1) << 16);   // next_arr patched below
            rec.hot.z = (unsigned)it->second | (tsm << 16);
            rec.hot.w = ecm | (edm << 16);
            rec.cap0 = b->s_cap0[r];
            {   // EV.calculate_max_energy_with_AFAP(cs.get_max_power())   ev.py:407-440, ev_charger.py:251-252,279
                const CsStatic &cs = h->cs_h[b->s_loc[r]];
                const double max_cs_power = cs.imax * (cs.veff[1]) * std::sqrt((double)cs.phases) / 1000.0;
                const double max_power = std::fabs(max_cs_power) > std::fabs(sp.pmax_ac) ? sp.pmax_ac : max_cs_power;
                double eff = b->s_eta_c[r];
                if (b->s_lut[r] >= 0) {
                    const double *l = b->luts_c + (b->lut_off[i] + b->s_lut[r]) * lut_len;
                    double m = 0; for (int q = 0; q < lut_len; ++q) m = std::max(m, l[q]);
                    eff = m / 100.0;
                }
                double afap = b->s_cap0[r];
                for (int q = b->s_t_arr[r]; q < b->s_t_dep[r] + 1; ++q) {
                    afap += max_power * eff * (double)h->dims.timescale / 60.0;
                    afap = std::ceil(afap * 100.0) / 100.0;
                    if (afap > sp.B) { afap = sp.B; break; }
                }
                rec.afap = afap;
                Lmax = std::max(Lmax, std::min(td, T) - ta + 3);
            }
            if (last_of_port[pl.port] >= 0) {
                SessRec &prev = sess[((size_t)i * P + pl.port) * Smax + last_of_port[pl.port]];
                prev.hot.y = (prev.hot.y & 0xFFFF0000u) | ((unsigned)ta & 0xFFFFu);
            }
            last_of_port[pl.port] = k;
            if (b->s_t_arr[r] <= T) {           // placed[] is arrival-sorted: the list comes out bucketed by step
                arr_list_h.push_back((unsigned)pl.port | ((unsigned)k << 16));
                arr_off_h[(size_t)i * (T + 2) + b->s_t_arr[r] + 1] += 1;
            }
        }
        int *ao = arr_off_h.data() + (size_t)i * (T + 2);        // ao[q + 1] holds the count of step q: prefix-sum into offsets
        int acc = (int)arr_base;
        ao[0] = acc;
        for (int q = 1; q <= T + 1; ++q) { acc += ao[q]; ao[q] = acc; }
    }
    // potential contribution per (spec, charger class)          utils.py:772-777
    std::vector<double> pot_kw(specs.size() * h->n_cls);
    {
        std::vector<int> cls_ph(h->n_cls, 3);
        for (const CsStatic &s : h->cs_h) cls_ph[s.cls] = s.phases;
        for (size_t q = 0; q < specs.size(); ++q)
            for (int k = 0; k < h->n_cls; ++k) {
                const int ph = std::min(cls_ph[k], specs[q].ev_phases);
                const double sv = h->cls_veff[k][ph];                       // sqrt(phases)*voltage
                const double ev_current = specs[q].pmax_ac * 1000.0 / sv;
                const double current = std::min(h->cls_imax[k], ev_current);
                pot_kw[q * h->n_cls + k] = sv * current / 1000.0;
            }
    }
    // ---- time series -----------------------------------------------------------------------------
    std::vector<EnvT> env_t((size_t)S * T);
    std::vector<TrT> tr_t((size_t)S * T * Tr);
    std::vector<float> trA((size_t)S * Tr * T), trF((size_t)S * Tr * T), tr_limit((size_t)S * Tr);
    std::vector<DrEv> dr((size_t)S * Tr * n_dr, DrEv{0, 0, 0.f});
    std::vector<uint8_t> dr_count((size_t)S * Tr, 0);
    for (int i = 0; i < S; ++i) {
        for (int t = 0; t < T; ++t) {
            EnvT &e = env_t[(size_t)i * T + t];
            e.cp = b->charge_price[(size_t)i * T + t]; e.dp = b->discharge_price[(size_t)i * T + t];
            e.setpoint = b->setpoint[(size_t)i * T + t];
            const int *ao = arr_off_h.data() + (size_t)i * (T + 2);      // the step that starts at t spawns the arrivals of t + 1
            e.arr0 = ao[t + 1]; e.n_arr = ao[t + 2] - ao[t + 1];
        }
        for (int k = 0; k < Tr; ++k) {
            const size_t base = ((size_t)i * Tr + k) * T;
            double limit = b->tr_max_power[base];
            for (int t = 0; t < T; ++t) {
                TrT &x = tr_t[((size_t)i * T + t) * Tr + k];
                x.infl = b->tr_infl[base + t]; x.solar = b->tr_solar[base + t];
                x.maxp = b->tr_max_power[base + t]; x.minp = b->tr_min_power[base + t];
                trA[base + t] = (float)(b->tr_infl[base + t] - b->tr_solar[base + t]);
                trF[base + t] = (float)(b->tr_load_fc[base + t] - b->tr_pv_fc[base + t]);
                limit = std::max(limit, x.maxp);                            // max(self.max_power) transformer.py:150
            }
            tr_limit[(size_t)i * Tr + k] = (float)limit;
            const int nd = b->dr_count ? b->dr_count[(size_t)i * Tr + k] : 0;
            if (nd > n_dr || nd > 255) return h->fail(EV2B_E_SCENARIO, "scenario %d: dr_count > n_dr", i);
            dr_count[(size_t)i * Tr + k] = (uint8_t)nd;
            for (int q = 0; q < nd; ++q) {
                const size_t di = ((size_t)i * Tr + k) * b->n_dr + q;
                DrEv &d = dr[((size_t)i * Tr + k) * n_dr + q];
                d.start = (int16_t)std::max(-32000, std::min(32000, b->dr_start[di]));
                d.end = (int16_t)std::max(-32000, std::min(32000, b->dr_end[di]));
                d.value = (float)(limit - limit * b->dr_cap[di] / 100.0);   // transformer.py:158-159
            }
        }
    }
    {
        const size_t nd = (size_t)S * (T + 1) * 3;
        std::vector<double> df(nd, 0.0);
        if (b->date_feat) std::copy(b->date_feat, b->date_feat + nd, df.begin());
        CUDA_TRY(h, h->date_feat.upload(df));
        if (h->n_bus > 0) {
            if (!b->grid_active || !b->grid_reactive) return h->fail(EV2B_E_SCENARIO, "grid handle: scenarios need grid_active / grid_reactive");
            const size_t ng = (size_t)S * (T + 1) * h->n_bus;
            std::vector<double> ga(b->grid_active, b->grid_active + ng), gr(b->grid_reactive, b->grid_reactive + ng);
            CUDA_TRY(h, h->grid_act.upload(ga)); CUDA_TRY(h, h->grid_rea.upload(gr));
        }
    }
    if (luts_c.empty()) { luts_c.assign(lut_len, 1.0); luts_d.assign(lut_len, 1.0); }
    if (specs.empty()) { specs.push_back(EvSpec{}); pot_kw.assign(h->n_cls, 0.0); }
    CUDA_TRY(h, h->env_t.upload(env_t)); CUDA_TRY(h, h->tr_t.upload(tr_t)); CUDA_TRY(h, h->sess.upload(sess));
    CUDA_TRY(h, h->spec.upload(specs)); CUDA_TRY(h, h->luts_c.upload(luts_c)); CUDA_TRY(h, h->luts_d.upload(luts_d));
    CUDA_TRY(h, h->pot_kw.upload(pot_kw)); CUDA_TRY(h, h->trA.upload(trA)); CUDA_TRY(h, h->trF.upload(trF));
    CUDA_TRY(h, h->tr_limit.upload(tr_limit)); CUDA_TRY(h, h->dr.upload(dr)); CUDA_TRY(h, h->dr_count.upload(dr_count));
    if (h->evl) {
        if (arr_list_h.empty()) arr_list_h.push_back(0u);
        if (h->spawn_ready) arr_list_h.resize(std::max(arr_list_h.size(), (size_t)S * P * Smax), 0u);   // a region per scenario
        CUDA_TRY(h, h->arr_list.upload(arr_list_h));
        h->list_valid = true;            // every env reads as done until it is reset, and reset empties its list
    }
    h->S = S; h->Smax = Smax; h->n_dr = n_dr; h->lut_len = lut_len;
    h->last_obs = nullptr;
    h->spawn_specs_live = false;
    if (h->spawn_ready && (h->dims.flags & EV2B_F_STATS)) Lmax = T + 2;   // sampled sessions can be as long as the episode
    if (h->dims.flags & EV2B_F_STATS) {
        h->L = std::max(2, std::min(Lmax, T + 2));
        const size_t EP = (size_t)h->E * P, EC = (size_t)h->E * C;
        CUDA_TRY(h, h->st_soc_sum.alloc(EP)); CUDA_TRY(h, h->st_abs_e.alloc(EP)); CUDA_TRY(h, h->st_cnt.alloc(EP));
        CUDA_TRY(h, h->st_nfin.alloc(EP)); CUDA_TRY(h, h->st_act.alloc(EP * h->L)); CUDA_TRY(h, h->st_r.alloc(EP * Smax));
        CUDA_TRY(h, h->cs_sat_sum.alloc(EC)); CUDA_TRY(h, h->cs_dcal.alloc(EC)); CUDA_TRY(h, h->cs_dcyc.alloc(EC));
        CUDA_TRY(h, h->cs_served.alloc(EC)); CUDA_TRY(h, h->cs_em.alloc(EC));
        CUDA_TRY(h, cudaMemset(h->st_cnt.p, 0, EP * sizeof(int))); CUDA_TRY(h, cudaMemset(h->st_nfin.p, 0, EP * sizeof(int)));
        CUDA_TRY(h, cudaMemset(h->cs_served.p, 0, EC * sizeof(int))); CUDA_TRY(h, cudaMemset(h->cs_em.p, 0, EC * sizeof(int)));
    }
    h->obs_static.release();
    if (h->W > 0) {   // precompute the (scenario, time)-only observation values when the table is small enough
        const size_t n = (size_t)S * (T + 1) * h->W;
        if (n * sizeof(float) <= ((size_t)1 << 31)) {
            DevBuf<float> tab;
            CUDA_TRY(h, tab.alloc(n));
            Params p = h->params();
            p.obs_static = nullptr;
            EV2B_LAUNCH(obs_static_kernel, (unsigned)std::min<size_t>(1024, (n + 255) / 256), 256, 0, (cudaStream_t)0, p, tab.p);
            CUDA_TRY(h, cudaGetLastError());
            CUDA_TRY(h, cudaDeviceSynchronize());
            h->obs_static.p = tab.p; h->obs_static.n = tab.n; tab.p = nullptr; tab.n = 0;
            h->launches += 1;
        }
    }
    {   // a new bank invalidates every running episode
        std::vector<int> st(h->E, h->T);
        CUDA_TRY(h, cudaMemcpy(h->env_step.p, st.data(), st.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    return EV2B_OK;
}

static int launch_reset(ev2b_handle *h, int lo, int hi, const int *scn_dev, int mode, float *obs0, cudaStream_t st) {
    const Params p = h->params();
    const size_t n = (size_t)(hi - lo) * h->P;
    const int blk = 256;
    EV2B_LAUNCH(reset_ports_kernel, (unsigned)((n + blk - 1) / blk), blk, 0, st, p, lo, hi, scn_dev, mode);
    EV2B_LAUNCH(reset_envs_kernel, hi - lo, 128, 0, st, p, lo, hi, scn_dev, mode, obs0);
    h->launches += 2;
    CUDA_TRY(h, cudaGetLastError());
    return EV2B_OK;
}

int ev2b_reset(ev2b_handle *h, int env_lo, int env_hi, const int32_t *scn_ids, float *obs0, void *stream) {
    if (!h) return EV2B_E_ARG;
    if (h->S == 0) return h->fail(EV2B_E_STATE, "reset: no scenario bank loaded");
    if (env_lo < 0 || env_hi > h->E || env_lo >= env_hi) return h->fail(EV2B_E_ARG, "reset: bad env range [%d,%d)", env_lo, env_hi);
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int *scn_dev = nullptr;
    if (scn_ids) {
        for (int i = 0; i < env_hi - env_lo; ++i)
            if (scn_ids[i] < 0 || scn_ids[i] >= h->S) return h->fail(EV2B_E_ARG, "reset: scenario id %d out of range", scn_ids[i]);
        if (h->st_scn.n < (size_t)h->E) CUDA_TRY(h, h->st_scn.alloc(h->E));
        CUDA_TRY(h, cudaMemcpyAsync(h->st_scn.p, scn_ids, sizeof(int) * (env_hi - env_lo), cudaMemcpyHostToDevice, st));
        CUDA_TRY(h, cudaStreamSynchronize(st));   // scn_ids is a caller-owned (possibly pageable) host buffer
        scn_dev = h->st_scn.p;
    }
    if (obs0 == nullptr || obs0 != h->last_obs) h->last_obs = nullptr;   // rows of this buffer are not all current
    if (obs0 != nullptr && env_lo == 0 && env_hi == h->E) h->last_obs = obs0;
    return launch_reset(h, env_lo, env_hi, scn_dev, 0, obs0, st);
}

int ev2b_reset_done(ev2b_handle *h, float *obs0, void *stream) {
    if (!h) return EV2B_E_ARG;
    if (h->S == 0) return h->fail(EV2B_E_STATE, "reset_done: no scenario bank loaded");
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (obs0 == nullptr || obs0 != h->last_obs) h->last_obs = nullptr;
    return launch_reset(h, 0, h->E, nullptr, 1, obs0, (cudaStream_t)stream);
}

int ev2b_step(ev2b_handle *h, const void *actions, int action_dtype, const ev2b_step_out *out, void *stream);

// Launches the step kernel for envs [lo, hi) on `st` (used by ev2b_step and by the chunked host path).
struct AgentCfg { int kind = EV2B_AGENT_EXTERNAL; uint64_t seed = 0; double low = -1.0; };

static int step_range(ev2b_handle *h, const void *actions, int action_dtype, const ev2b_step_out *out, int lo, int hi,
                      int obs_full, cudaStream_t st, const AgentCfg &ag = AgentCfg()) {
    Params p = h->params();
    p.agent_kind = ag.kind; p.agent_seed_lo = (unsigned)ag.seed; p.agent_seed_hi = (unsigned)(ag.seed >> 32);
    p.action_low = ag.low;
    p.actions = actions;
    if (out) p.out = *out;
    p.obs_full = obs_full;
    p.mask_full = (p.out.action_mask && p.out.action_mask != h->last_mask) ? 1 : 0;
    // after this launch the mask buffer is current only if the launch wrote it for every env (both kernels do)
    h->last_mask = (lo == 0 && hi == h->E) ? p.out.action_mask : nullptr;
    p.env0 = lo; p.env_end = hi;
    cudaError_t e;
    if (action_dtype != EV2B_F32 && action_dtype != EV2B_F64) return h->fail(EV2B_E_ARG, "step: unknown action dtype %d", action_dtype);
    if (evl_covers(h, out)) {            // callers ran ensure_list() on a stream this launch is ordered after
        if (!h->list_valid) return h->fail(EV2B_E_STATE, "step: connected-EV list is stale (internal error)");
        e = action_dtype == EV2B_F32 ? launch_evl<float>(h, p, st) : launch_evl<double>(h, p, st);
        h->launches_by[1] += 1;
    } else {
        if (h->evl) h->list_valid = false;
        h->launches_by[0] += 1;
        e = action_dtype == EV2B_F32 ? launch_step<float>(h, p, st) : launch_step<double>(h, p, st);
    }
    if (e != cudaSuccess) return h->fail(EV2B_E_CUDA, "step launch: %s", cudaGetErrorString(e));
    h->launches += 1;
    return EV2B_OK;
}

int ev2b_step(ev2b_handle *h, const void *actions, int action_dtype, const ev2b_step_out *out, void *stream) {
    if (!h || !actions) return h ? h->fail(EV2B_E_ARG, "step: null actions") : EV2B_E_ARG;
    if (h->S == 0) return h->fail(EV2B_E_STATE, "step: no scenario bank loaded");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const float *obs = out ? out->obs : nullptr;
    const int obs_full = (obs != h->last_obs) ? 1 : 0;
    h->last_obs = obs;
    if (evl_covers(h, out)) { const int rc = ensure_list(h, (cudaStream_t)stream); if (rc != EV2B_OK) return rc; }
    return step_range(h, actions, action_dtype, out, 0, h->E, obs_full, (cudaStream_t)stream);
}

int ev2b_agent_actions(ev2b_handle *h, int agent_kind, double *actions_out, void *stream) {
    if (!h || !actions_out) return h ? h->fail(EV2B_E_ARG, "agent_actions: null output") : EV2B_E_ARG;
    if (h->S == 0) return h->fail(EV2B_E_STATE, "agent_actions: no scenario bank loaded");
    if (agent_kind != EV2B_AGENT_AFAP && agent_kind != EV2B_AGENT_ZERO && agent_kind != EV2B_AGENT_ROUNDROBIN &&
        agent_kind != EV2B_AGENT_CALAP)
        return h->fail(EV2B_E_ARG, "agent_actions: agent %d has no action tensor (AFAP, ZERO, ROUNDROBIN, CALAP)", agent_kind);
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (agent_kind == EV2B_AGENT_ROUNDROBIN && !h->rr_key.p) {      // first use: every env starts with an empty queue
        std::vector<int> keys((size_t)h->E * h->P, kRrAbsent), fb((size_t)h->E * 2);
        for (int e = 0; e < h->E; ++e) { fb[2 * e] = 0; fb[2 * e + 1] = 1; }
        CUDA_TRY(h, h->rr_key.upload(keys));
        CUDA_TRY(h, h->rr_fb.upload(fb));
        CUDA_TRY(h, cudaStreamSynchronize(st));
    }
    const int thr = std::min(kMaxThreads, std::max(32, (h->P + 31) / 32 * 32));
    const size_t sm = (size_t)h->P * 5 + 16;
    CUDA_TRY(h, opt_in_smem(h, agent_kernel, sm));
    EV2B_LAUNCH(agent_kernel, h->E, thr, sm, st, h->params(), agent_kind, actions_out);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return EV2B_OK;
}

int ev2b_step_k(ev2b_handle *h, int k, int agent_kind, const void *actions_k, int action_dtype, uint64_t seed,
                double action_low, int auto_reset, const ev2b_step_out *out, void *stream) {
    if (!h) return EV2B_E_ARG;
    if (h->S == 0) return h->fail(EV2B_E_STATE, "step_k: no scenario bank loaded");
    if (k < 1) return h->fail(EV2B_E_ARG, "step_k: k must be >= 1");
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (agent_kind < EV2B_AGENT_EXTERNAL || agent_kind > EV2B_AGENT_CALAP) return h->fail(EV2B_E_ARG, "step_k: unknown agent %d", agent_kind);
    const bool tensor_agent = agent_kind == EV2B_AGENT_ROUNDROBIN || agent_kind == EV2B_AGENT_CALAP;   // needs the whole env's state
    if (tensor_agent && h->agent_act.n < (size_t)h->E * h->P) CUDA_TRY(h, h->agent_act.alloc((size_t)h->E * h->P));
    if (agent_kind == EV2B_AGENT_EXTERNAL && !actions_k) return h->fail(EV2B_E_ARG, "step_k: EXTERNAL agent needs actions_k");
    AgentCfg ag; ag.kind = agent_kind; ag.seed = seed; ag.low = action_low;
    const size_t stride = (size_t)h->E * h->P * (action_dtype == EV2B_F64 ? 8 : 4);
    float *obs = out ? out->obs : nullptr;
    // Agents that need nothing but the env's own step (an action tensor, AFAP, ZERO, UNIFORM) on the lean event-driven
    // kernel: ONE launch advances every env k steps (evl_step_kernel<..., KSTEP = true>); EV2B_STEP_K=loop keeps the
    // launch-per-step path (A/B, tests).
    const bool heavy_out = out && (out->dep_sat || out->dep_cap || out->port_energy || out->node_voltage);
    const char *kv = getenv("EV2B_STEP_K");
    if (!tensor_agent && k > 1 && evl_covers(h, out) && !h->evl_heavy_layout() && !heavy_out && !(kv && !strcmp(kv, "loop"))) {
        if (action_dtype != EV2B_F32 && action_dtype != EV2B_F64) return h->fail(EV2B_E_ARG, "step_k: unknown action dtype %d", action_dtype);
        int rc = ensure_list(h, (cudaStream_t)stream);
        if (rc != EV2B_OK) return rc;
        Params p = h->params();
        p.agent_kind = ag.kind; p.agent_seed_lo = (unsigned)ag.seed; p.agent_seed_hi = (unsigned)(ag.seed >> 32);
        p.action_low = ag.low;
        p.actions = agent_kind == EV2B_AGENT_EXTERNAL ? actions_k : (const void *)h->hot.p;
        if (out) p.out = *out;
        p.obs_full = (obs != h->last_obs) ? 1 : 0;
        p.mask_full = (p.out.action_mask && p.out.action_mask != h->last_mask) ? 1 : 0;
        h->last_obs = obs; h->last_mask = p.out.action_mask;
        p.k_steps = k; p.auto_reset = auto_reset ? 1 : 0;
        const cudaError_t e = action_dtype == EV2B_F32 ? launch_evl<float>(h, p, (cudaStream_t)stream, true)
                                                       : launch_evl<double>(h, p, (cudaStream_t)stream, true);
        if (e != cudaSuccess) return h->fail(EV2B_E_CUDA, "step_k launch: %s", cudaGetErrorString(e));
        h->launches += 1; h->launches_by[1] += 1;
        return EV2B_OK;
    }
    for (int i = 0; i < k; ++i) {
        const int obs_full = (obs != h->last_obs) ? 1 : 0;
        h->last_obs = obs;
        const void *a = agent_kind == EV2B_AGENT_EXTERNAL ? (const void *)((const unsigned char *)actions_k + stride * i) : (const void *)h->hot.p;
        int rc;
        if (evl_covers(h, out)) { rc = ensure_list(h, (cudaStream_t)stream); if (rc != EV2B_OK) return rc; }
        if (tensor_agent) {
            rc = ev2b_agent_actions(h, agent_kind, h->agent_act.p, stream);
            if (rc != EV2B_OK) return rc;
            rc = step_range(h, h->agent_act.p, EV2B_F64, out, 0, h->E, obs_full, (cudaStream_t)stream);
        } else {
            rc = step_range(h, a, action_dtype, out, 0, h->E, obs_full, (cudaStream_t)stream, ag);
        }
        if (rc != EV2B_OK) return rc;
        if (auto_reset) {
            rc = ev2b_reset_done(h, obs, stream);
            if (rc != EV2B_OK) return rc;
        }
    }
    return EV2B_OK;
}

int ev2b_step_host(ev2b_handle *h, const void *actions_host, int action_dtype, double *reward_host,
                   uint32_t *status_host, float *obs_host, void *stream) {
    if (!h || !actions_host) return h ? h->fail(EV2B_E_ARG, "step_host: null actions") : EV2B_E_ARG;
    if (h->S == 0) return h->fail(EV2B_E_STATE, "step_host: no scenario bank loaded");
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t user = (cudaStream_t)stream;
    const size_t esz = action_dtype == EV2B_F64 ? 8 : 4;
    const size_t nbytes = (size_t)h->E * h->P * esz;
    if (h->st_actions.n < nbytes) CUDA_TRY(h, h->st_actions.alloc(nbytes));
    if (h->st_reward.n < (size_t)h->E) { CUDA_TRY(h, h->st_reward.alloc(h->E)); CUDA_TRY(h, h->st_status.alloc(h->E)); }
    const bool want_obs = obs_host && h->D > 0;
    if (want_obs && h->st_obs.n < (size_t)h->E * h->D) CUDA_TRY(h, h->st_obs.alloc((size_t)h->E * h->D));
    if (!h->start_ev) {
        if (const char *ov = getenv("EV2B_HOST_CHUNKS")) h->n_chunks = std::min(ev2b_handle::kChunks, std::max(1, atoi(ov)));
        CUDA_TRY(h, cudaEventCreateWithFlags(&h->start_ev, cudaEventDisableTiming));
        for (int i = 0; i < ev2b_handle::kChunks; ++i) {
            CUDA_TRY(h, cudaStreamCreateWithFlags(&h->chunk_stream[i], cudaStreamNonBlocking));
            CUDA_TRY(h, cudaEventCreateWithFlags(&h->chunk_ev[i], cudaEventDisableTiming));
        }
    }
    ev2b_step_out out{};
    out.reward = h->st_reward.p; out.status = h->st_status.p;
    out.obs = want_obs ? h->st_obs.p : nullptr;
    const int obs_full = (out.obs != h->last_obs) ? 1 : 0;
    h->last_obs = out.obs;
    if (evl_covers(h, &out)) { const int rc = ensure_list(h, user); if (rc != EV2B_OK) return rc; }   // before start_ev: every chunk stream waits on it
    // PCIe is full duplex and the copy engines run beside the SMs: split the env range into chunks, each on its
    // own stream (H2D actions -> kernel -> D2H results), so chunk i's download overlaps chunk i+1's upload/compute.
    const int per = ((h->E + h->n_chunks - 1) / h->n_chunks + h->EPB - 1) / h->EPB * h->EPB;
    const unsigned char *ah = static_cast<const unsigned char *>(actions_host);
    auto issue = [&](cudaStream_t root) -> int {
        CUDA_TRY(h, cudaEventRecord(h->start_ev, root));
        for (int c = 0; c < h->n_chunks; ++c) {
            const int lo = c * per, hi = std::min(h->E, lo + per);
            if (lo >= hi) break;
            cudaStream_t st = h->chunk_stream[c];
            CUDA_TRY(h, cudaStreamWaitEvent(st, h->start_ev, 0));
            const size_t aoff = (size_t)lo * h->P * esz, abytes = (size_t)(hi - lo) * h->P * esz;
            CUDA_TRY(h, cudaMemcpyAsync(h->st_actions.p + aoff, ah + aoff, abytes, cudaMemcpyHostToDevice, st));
            int rc = step_range(h, h->st_actions.p, action_dtype, &out, lo, hi, obs_full, st);
            if (rc != EV2B_OK) return rc;
            if (reward_host) CUDA_TRY(h, cudaMemcpyAsync(reward_host + lo, h->st_reward.p + lo, sizeof(double) * (hi - lo), cudaMemcpyDeviceToHost, st));
            if (status_host) CUDA_TRY(h, cudaMemcpyAsync(status_host + lo, h->st_status.p + lo, sizeof(uint32_t) * (hi - lo), cudaMemcpyDeviceToHost, st));
            if (want_obs) CUDA_TRY(h, cudaMemcpyAsync(obs_host + (size_t)lo * h->D, h->st_obs.p + (size_t)lo * h->D,
                                                      sizeof(float) * (size_t)(hi - lo) * h->D, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(h, cudaEventRecord(h->chunk_ev[c], st));
            CUDA_TRY(h, cudaStreamWaitEvent(root, h->chunk_ev[c], 0));
        }
        return EV2B_OK;
    };
    const int rc = issue(user);
    if (rc != EV2B_OK) return rc;
    CUDA_TRY(h, cudaStreamSynchronize(user));
    return EV2B_OK;
}

int ev2b_episode_stats(ev2b_handle *h, double *out, void *stream) {
    if (!h || !out) return h ? h->fail(EV2B_E_ARG, "episode_stats: null output") : EV2B_E_ARG;
    if (!(h->dims.flags & EV2B_F_STATS)) return h->fail(EV2B_E_STATE, "episode_stats: handle was created without EV2B_F_STATS");
    if (h->S == 0) return h->fail(EV2B_E_STATE, "episode_stats: no scenario bank loaded");
    CUDA_TRY(h, cudaSetDevice(h->device));
    EV2B_LAUNCH(episode_stats_kernel, h->E, 32, 0, (cudaStream_t)stream, h->params(), out);
    h->launches += 1;
    CUDA_TRY(h, cudaGetLastError());
    return EV2B_OK;
}

#ifdef EV2B_SIMT_EMU
// emulator builds only (tests/test_emu_kernels.py): the connected-EV list of env e
int ev2b_debug_list(ev2b_handle *h, int e, uint16_t *out) {
    if (!h || !h->evl || e < 0 || e >= h->E) return -1;
    const int n = h->occ_n.p[e];
    memcpy(out, h->occ_list.p + (size_t)e * h->P, sizeof(uint16_t) * (size_t)n);
    return n;
}
#endif

int ev2b_set_spawn_tables(ev2b_handle *h, const ev2b_spawn_tables *t) {
    if (!h || !t) return h ? h->fail(EV2B_E_ARG, "set_spawn_tables: null tables") : EV2B_E_ARG;
    if (h->S != 0) return h->fail(EV2B_E_STATE, "set_spawn_tables: call before ev2b_load_scenarios");
    if (!h->evl) return h->fail(EV2B_E_STATE, "set_spawn_tables: the device sampler needs the event-driven kernel (EV2B_KERNEL=evlist)");
    if (t->n_models < 1 || t->n_models > 65535 || t->min_stay_steps < 0) return h->fail(EV2B_E_ARG, "set_spawn_tables: bad sizes");
    if (((size_t)(h->T + 2) * ((h->P + 31) / 32) + h->T + 3) * 4 > 200 * 1024)
        return h->fail(EV2B_E_LIMIT, "set_spawn_tables: too many ports for the arrival-schedule kernel");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int M = t->n_models, LL = 101;
    auto up = [&](DevBuf<double> &b, const double *src, size_t n) { return b.upload(std::vector<double>(src, src + n)); };
    CUDA_TRY(h, up(h->sp_arr_week, t->arrival_week, 96)); CUDA_TRY(h, up(h->sp_arr_weekend, t->arrival_weekend, 96));
    CUDA_TRY(h, up(h->sp_req, t->req_energy_mean, 48)); CUDA_TRY(h, up(h->sp_stay, t->stay_mean, 48));
    std::vector<double> cdf(M); double acc = 0, tot = 0;
    for (int i = 0; i < M; ++i) tot += t->model_prob[i];
    for (int i = 0; i < M; ++i) { acc += t->model_prob[i] / tot; cdf[i] = acc; }
    CUDA_TRY(h, h->sp_cdf.upload(cdf)); CUDA_TRY(h, up(h->sp_B, t->model_B, M));
    CUDA_TRY(h, h->sp_lut.upload(std::vector<int>(t->model_lut, t->model_lut + M)));
    std::vector<double> luts(t->n_luts > 0 ? (size_t)t->n_luts * LL : (size_t)LL, 1.0);
    if (t->n_luts > 0) std::copy(t->luts, t->luts + (size_t)t->n_luts * LL, luts.begin());
    CUDA_TRY(h, h->sp_luts.upload(luts));
    // the EV-model spec table (what ev2b_load_scenarios de-duplicates from a bank's sessions)   ev.py:45-113, utils.py:298-345
    std::vector<EvSpec> specs(M);
    for (int i = 0; i < M; ++i) {
        EvSpec sp{};
        sp.B = t->model_B[i]; sp.pmax_ac = t->model_pmax_ac[i]; sp.pmin_ac = t->model_pmin_ac[i];
        sp.pmax_dis = t->model_pmax_dis[i]; sp.pmin_dis = t->model_pmin_dis[i];
        sp.bmin = t->min_battery_capacity;
        sp.bmin_em = t->min_emergency_battery_capacity > sp.B ? 0.7 * sp.B : t->min_emergency_battery_capacity;   // utils.py:268-271
        sp.desired = t->desired_frac * sp.B; sp.mult = t->ts_multiplier;
        sp.ts = t->homog_ts; sp.eta_c = t->homog_eta_c; sp.eta_d = t->homog_eta_d;
        sp.ev_phases = t->model_phases[i]; sp.lut = t->model_lut[i];
        if (sp.ev_phases < 1 || sp.ev_phases > 3 || sp.lut >= t->n_luts) return h->fail(EV2B_E_ARG, "set_spawn_tables: bad model %d", i);
        sp.rB = 1.0 / sp.B;
        specs[i] = sp;
    }
    CUDA_TRY(h, h->sp_spec.upload(specs));
    std::vector<double> pot_kw((size_t)M * h->n_cls);
    {
        std::vector<int> cls_ph(h->n_cls, 3);
        for (const CsStatic &c : h->cs_h) cls_ph[c.cls] = c.phases;
        for (int q = 0; q < M; ++q)
            for (int k = 0; k < h->n_cls; ++k) {                           // utils.py:772-777
                const int ph = std::min(cls_ph[k], specs[q].ev_phases);
                const double sv = h->cls_veff[k][ph];
                pot_kw[(size_t)q * h->n_cls + k] = sv * std::min(h->cls_imax[k], specs[q].pmax_ac * 1000.0 / sv) / 1000.0;
            }
    }
    CUDA_TRY(h, h->sp_pot_kw.upload(pot_kw));
    SpawnParams &sp = h->spawn_p;
    sp = SpawnParams{};
    sp.arrival_week = h->sp_arr_week.p; sp.arrival_weekend = h->sp_arr_weekend.p; sp.req_energy_mean = h->sp_req.p;
    sp.stay_mean = h->sp_stay.p; sp.model_cdf = h->sp_cdf.p; sp.model_B = h->sp_B.p; sp.model_lut = h->sp_lut.p;
    sp.M = M; sp.workplace = t->workplace; sp.heterogeneous = t->heterogeneous; sp.empty_ports_at_end = t->empty_ports_at_end;
    sp.min_stay_steps = t->min_stay_steps; sp.timescale = h->dims.timescale;
    sp.spawn_multiplier = t->spawn_multiplier; sp.desired_frac = t->desired_frac; sp.min_battery_capacity = t->min_battery_capacity;
    h->spawn_setpoints = t->power_setpoint_enabled != 0;
    sp.setpoint_mult = 100.0 + t->power_setpoint_flexibility;
    {   // charging_stations[0].get_min_charge_power() / get_max_power()   ev_charger.py:251-255
        const CsStatic &c0 = h->cs_h[0];
        sp.min_cs_power = c0.imin * c0.veff[1] * std::sqrt((double)c0.phases) / 1000.0;
        sp.max_cs_power = c0.imax * c0.veff[1] * std::sqrt((double)c0.phases) / 1000.0;
    }
    sp.median_window = 5 * std::max(1, 15 / h->dims.timescale);
    sp.setpoint_threads = 128;                       // 2 * threads * T doubles of shared memory must fit
    while (sp.setpoint_threads > 1 && (size_t)(2 * sp.setpoint_threads + 2) * h->T * 8 > 200 * 1024) sp.setpoint_threads /= 2;
    if (h->spawn_setpoints && (size_t)(2 * sp.setpoint_threads + 2) * h->T * 8 > 200 * 1024)
        return h->fail(EV2B_E_LIMIT, "set_spawn_tables: episode too long for the setpoint generator");
    unsigned k = 0xFFFFu;
    sp.homog_ts_milli = milli(t->homog_ts, &k) ? k : 0xFFFFu;
    sp.homog_eta_c_milli = milli(t->homog_eta_c, &k) ? k : 0xFFFFu;
    sp.homog_eta_d_milli = milli(t->homog_eta_d, &k) ? k : 0xFFFFu;
    // a session holds its port for at least min_stay + 2 steps and the port then rests for 2 more   utils.py:534-536, 551-552
    h->spawn_smax = std::min(255, h->T / std::max(1, t->min_stay_steps + 4) + 2);
    h->spawn_M = M;
    h->spawn_ready = true;
    return EV2B_OK;
}

int ev2b_resample_sessions(ev2b_handle *h, uint64_t seed, const int32_t *start, void *stream) {
    if (!h || !start) return h ? h->fail(EV2B_E_ARG, "resample_sessions: null start dates") : EV2B_E_ARG;
    if (!h->spawn_ready) return h->fail(EV2B_E_STATE, "resample_sessions: ev2b_set_spawn_tables was not called");
    if (h->S == 0) return h->fail(EV2B_E_STATE, "resample_sessions: no scenario bank loaded");
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t SP = (size_t)h->S * h->P;
    for (int i = 0; i < h->S; ++i)
        if (start[3 * i] < 0 || start[3 * i] > 6 || start[3 * i + 1] < 0 || start[3 * i + 1] > 23 || start[3 * i + 2] < 0 || start[3 * i + 2] > 59)
            return h->fail(EV2B_E_ARG, "resample_sessions: bad start date of scenario %d", i);
    if (h->sp_raw.n < SP * h->Smax) { CUDA_TRY(h, h->sp_raw.alloc(SP * h->Smax)); CUDA_TRY(h, h->sp_raw_n.alloc(SP)); CUDA_TRY(h, h->sp_n_sess.alloc(h->S)); }
    if (h->sp_start.n < (size_t)3 * h->S) CUDA_TRY(h, h->sp_start.alloc((size_t)3 * h->S));
    CUDA_TRY(h, cudaMemcpyAsync(h->sp_start.p, start, sizeof(int) * 3 * h->S, cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));            // `start` is a caller-owned (possibly pageable) host buffer
    CUDA_TRY(h, cudaMemsetAsync(h->sp_n_sess.p, 0, sizeof(int) * h->S, st));
    if (!h->spawn_specs_live) {                        // from now on the bank's sessions name EV models, not the bank's own specs
        CUDA_TRY(h, cudaStreamSynchronize(st));
        CUDA_TRY(h, h->spec.alloc(h->sp_spec.n));
        CUDA_TRY(h, cudaMemcpy(h->spec.p, h->sp_spec.p, h->sp_spec.n * sizeof(EvSpec), cudaMemcpyDeviceToDevice));
        CUDA_TRY(h, h->pot_kw.alloc(h->sp_pot_kw.n));
        CUDA_TRY(h, cudaMemcpy(h->pot_kw.p, h->sp_pot_kw.p, h->sp_pot_kw.n * sizeof(double), cudaMemcpyDeviceToDevice));
        CUDA_TRY(h, h->luts_c.alloc(h->sp_luts.n)); CUDA_TRY(h, h->luts_d.alloc(h->sp_luts.n));
        CUDA_TRY(h, cudaMemcpy(h->luts_c.p, h->sp_luts.p, h->sp_luts.n * sizeof(double), cudaMemcpyDeviceToDevice));
        CUDA_TRY(h, cudaMemcpy(h->luts_d.p, h->sp_luts.p, h->sp_luts.n * sizeof(double), cudaMemcpyDeviceToDevice));
        h->lut_len = 101;
        h->spawn_specs_live = true;
    }
    SpawnParams sp = h->spawn_p;
    sp.start = h->sp_start.p; sp.seed_lo = (unsigned)seed; sp.seed_hi = (unsigned)(seed >> 32);
    sp.cap_per_scn = h->P * h->Smax;
    sp.raw = h->sp_raw.p; sp.raw_n = h->sp_raw_n.p; sp.sess = h->sess.p; sp.env_t = h->env_t.p; sp.arr_list = h->arr_list.p;
    sp.n_sess = h->sp_n_sess.p;
    const Params p = h->params();
    const int blk = 128;
    EV2B_LAUNCH(spawn_sessions_kernel, (unsigned)((SP + blk - 1) / blk), blk, 0, st, p, sp);
    EV2B_LAUNCH(spawn_assign_kernel, (unsigned)(((size_t)h->S * h->C + blk - 1) / blk), blk, 0, st, p, sp);
    const size_t sm = ((size_t)(h->T + 2) * ((h->P + 31) / 32) + h->T + 3) * 4;
    CUDA_TRY(h, opt_in_smem(h, spawn_schedule_kernel, sm));
    EV2B_LAUNCH(spawn_schedule_kernel, (unsigned)h->S, blk, sm, st, p, sp);
    if (h->spawn_setpoints) {                          // generate_power_setpoints from the new sessions  utils.py:664-757
        const size_t sm2 = (size_t)(2 * sp.setpoint_threads + 2) * h->T * 8;
        CUDA_TRY(h, opt_in_smem(h, spawn_setpoints_kernel, sm2));
        EV2B_LAUNCH(spawn_setpoints_kernel, (unsigned)h->S, sp.setpoint_threads, sm2, st, p, sp);
        h->launches += 1;
        if (h->obs_static.p && h->dims.state_kind == EV2B_STATE_V2G_GRID) {   // (the only state whose static part holds setpoints)
            Params q = p; q.obs_static = nullptr;
            const size_t n = (size_t)h->S * (h->T + 1) * h->W;
            EV2B_LAUNCH(obs_static_kernel, (unsigned)std::min<size_t>(1024, (n + 255) / 256), 256, 0, st, q, h->obs_static.p);
            h->launches += 1;
        }
    }
    EV2B_LAUNCH(spawn_invalidate_envs_kernel, (unsigned)((h->E + blk - 1) / blk), blk, 0, st, p);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 4;
    h->last_obs = nullptr; h->last_mask = nullptr;
    return EV2B_OK;
}

int ev2b_read_sessions(ev2b_handle *h, int scn, int cap, int32_t *port, int32_t *t_arr, int32_t *t_dep, int32_t *model,
                       double *cap0, double *ts, double *eta_c, double *eta_d) {
    if (!h) return EV2B_E_ARG;
    if (h->S == 0 || scn < 0 || scn >= h->S) return h->fail(EV2B_E_ARG, "read_sessions: scenario %d out of range", scn);
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaDeviceSynchronize());
    std::vector<SessRec> rec((size_t)h->P * h->Smax);
    CUDA_TRY(h, cudaMemcpy(rec.data(), h->sess.p + (size_t)scn * h->P * h->Smax, rec.size() * sizeof(SessRec), cudaMemcpyDeviceToHost));
    struct Row { int ta, port, k; };
    std::vector<Row> rows;
    for (int pp = 0; pp < h->P; ++pp)
        for (int k = 0; k < h->Smax; ++k) {
            const int ta = (int)(rec[(size_t)pp * h->Smax + k].hot.x & 0xFFFFu);
            if (ta == kNoArrival) break;
            rows.push_back({ta, pp, k});
        }
    std::stable_sort(rows.begin(), rows.end(), [](const Row &a, const Row &b) { return a.ta != b.ta ? a.ta < b.ta : a.port < b.port; });
    const double nan = std::nan("");
    int n = 0;
    for (const Row &r : rows) {
        if (n >= cap) break;
        const SessRec &x = rec[(size_t)r.port * h->Smax + r.k];
        if (port) port[n] = r.port;
        if (t_arr) t_arr[n] = r.ta;
        if (t_dep) t_dep[n] = (int)(int16_t)(x.hot.x >> 16);
        if (model) model[n] = (int)(x.hot.z & 0xFFFFu);
        if (cap0) cap0[n] = x.cap0;
        const unsigned tsm = x.hot.z >> 16, ecm = x.hot.w & 0xFFFFu, edm = x.hot.w >> 16;
        if (ts) ts[n] = tsm == 0xFFFFu ? nan : tsm / 1000.0;
        if (eta_c) eta_c[n] = (ecm == 0xFFFFu || ecm == 0) ? nan : ecm / 1000.0;
        if (eta_d) eta_d[n] = (edm == 0xFFFFu || edm == 0) ? nan : edm / 1000.0;
        ++n;
    }
    return (int)rows.size();
}

int ev2b_read_setpoints(ev2b_handle *h, int scn, double *out) {
    if (!h || !out) return EV2B_E_ARG;
    if (h->S == 0 || scn < 0 || scn >= h->S) return h->fail(EV2B_E_ARG, "read_setpoints: scenario %d out of range", scn);
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaDeviceSynchronize());
    std::vector<EnvT> et((size_t)h->T);
    CUDA_TRY(h, cudaMemcpy(et.data(), h->env_t.p + (size_t)scn * h->T, et.size() * sizeof(EnvT), cudaMemcpyDeviceToHost));
    for (int t = 0; t < h->T; ++t) out[t] = et[t].setpoint;
    return EV2B_OK;
}

int ev2b_state_view_get(ev2b_handle *h, ev2b_state_view *v) {
    if (!h || !v) return EV2B_E_ARG;
    v->n_envs = h->E; v->n_ports = h->P; v->n_chargers = h->C; v->n_transformers = h->Tr; v->obs_dim = h->D;
    v->n_kpi = EV2B_KPI_COUNT;
    v->port_cap = h->cap.p; v->port_exch = h->exch.p; v->port_hot = reinterpret_cast<uint32_t *>(h->hot.p);
    v->env_step = h->env_step.p; v->env_scn = h->env_scn.p; v->env_potential = h->env_pot.p;
    v->env_usage = h->env_usage.p; v->env_kpi = h->env_kpi.p;
    return EV2B_OK;
}

}  // extern "C"
