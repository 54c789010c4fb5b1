// Host side of libtmglow_b200: model description, parameter table, workspace planning and the
// launch sequences behind the C ABI (include/tmglow_b200.h).  The sequences restate
//   Encoder.forward            nn/tmGlow.py:104-129
//   LSTMFLowBlock.forward/rev  nn/modules/flowLSTMBlock.py:280-361
//   LSTMCFlowDecoder           nn/tmGlow.py:231-303
//   TMGlow.forward/reconstruct nn/tmGlow.py:378-467
// with every torch.cat / chunk / relu / pad tensor of the reference folded into the kernels.
#include <map>
#include <mutex>
#include <unordered_map>
#include <cstdarg>
#include <mutex>

#include "common.cuh"

namespace tmg {

std::atomic<int64_t> g_launches{0};
cudaError_t ensure_dyn_smem(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::unordered_map<const void*, std::vector<int>> done;     // kernel -> largest size set per device
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(mu);
  std::vector<int>& v = done[kernel];
  if ((int)v.size() <= dev) v.resize(dev + 1, -1);
  if (v[dev] >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) v[dev] = bytes;
  return e;
}   // process-wide: the backward runs on autograd's thread
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- per-class event timing
struct ProfRec { cudaEvent_t a, b; int tag; double flops, bytes; };
static thread_local bool g_prof_on = false;
static thread_local std::vector<ProfRec> g_prof;

ProfScope::ProfScope(cudaStream_t st_, int tag, double flops, double bytes) : st(st_), idx(-1) {
  if (!g_prof_on) return;
  ProfRec r{};
  r.tag = tag; r.flops = flops; r.bytes = bytes;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
  idx = (int)g_prof.size() - 1;
}
ProfScope::~ProfScope() {
  if (idx >= 0) cudaEventRecord(g_prof[idx].b, st);
}

struct ParamEntry { std::string name; int64_t off; int64_t numel; int ndim; int64_t dims[4]; };

struct ConvW {
  int64_t w_param = -1;   // OIHW in the parameter buffer
  int64_t b_param = -1;   // bias or -1
  int64_t w_pack = -1;    // tap-major in the packed buffer
  int64_t w_pack_tc = -1; // tcgen05 layout (hi/lo split) in the packed buffer, -1: CUDA-core path only
  int64_t w_pack_f16 = -1, inv_f16 = -1;   // fp16 hi/lo layout of conv3x3_f16.cu + inverse weight scale
  int f16_nch[3] = {0, 0, 0};              // channel split of the sources the fp16 weights were packed for
  int64_t w_pack_f16t = -1, inv_f16t = -1; // transposed / tap-flipped fp16 weights (data gradient, JOB_CONV_F16_T)
  int NPt = 0;                             // their MMA N (padded I)
  int O = 0, I = 0, OP = 0, NP = 0;
};

enum StepKind { STEP_UNNORMED = 0, STEP_PLAIN = 1, STEP_LSTM = 2 };

// TMG_NO_GATE2P=1 (read when the model is built): ConvLSTM gate conv through conv3x3_f16.cu instead of lstm_gate_f16.cu (A/B runs)
static inline bool gate2p_off() {
  static const bool off = [] { const char* e = getenv("TMG_NO_GATE2P"); return e && e[0] == '1'; }();
  return off;
}

// TMG_NO_MIXMMA=1 (read when the model is built): the 1x1 mix of the level-resident kernel on the CUDA cores (A/B runs)
static inline bool mixmma_off() {
  static const bool off = [] { const char* e = getenv("TMG_NO_MIXMMA"); return e && e[0] == '1'; }();
  return off;
}

static inline int mixmma_min_c() {       // TMG_MIXMMA_MINC=n (A/B runs): smallest C with the tensor-core mix
  static const int c = [] { const char* e = getenv("TMG_MIXMMA_MINC"); return e ? atoi(e) : 24; }();
  return c;
}

struct StepW {
  int kind = STEP_PLAIN;
  int64_t norm_w = -1, norm_b = -1;
  int64_t W = -1, Wi = -1;        // packed
  int64_t Wmx = -1;               // packed: W as fp16 hi/lo tensor-core operand (level-resident kernel, C >= 24)
  int const_idx = -1;             // index into step_const
  ConvW d1, d2, zc, gate, outc;
  int64_t zc_gain = -1;           // packed
  // fused tensor-core coupling kernel (coupling_tc.cu): packed weights and K-plane bookkeeping
  int64_t cpl_w1 = -1, cpl_w2 = -1, cpl_w3 = -1;
  int cpl_nch0 = 0, cpl_nch1 = 0, cpl_npad = 0;
  // fp16 fused step (flow_step_f16.cu)
  int64_t s2_wE = -1, s2_wZ = -1, s2_misc = -1;
  int s2_nch0 = 0, s2_nch1 = 0;
  int64_t s2c_wE = -1, s2c_wZ = -1;   // compact / tap-paired variant for the level-resident kernel (narrow levels)
  // LSTM step, one LF input shared by all samples: gate / output convolutions WITHOUT their conditioning rows (fp16 packing
  // over the sources [x1 | h]) + the conditioning rows as tap-major fp32 slices for the once-per-call hoisted tables
  ConvW gate_nc, outc_nc;
  int64_t gate_hw = -1, outc_hw = -1;
  int gate_hop = 0, outc_hop = 0;
  // two-pass gate kernel (lstm_gate_f16.cu, R = 64): weights over the sources [h | cond | x1] and [h | x1], gate rows in pass
  // order; gate2p_hw >= 0: gate_hw holds the conditioning slice in pass order (the hoisted table is then kept plane-transposed)
  ConvW gate2p, gate2p_nc;
  int64_t gate2p_hw = -1;
  // parameter offsets needed by the backward pass
  int64_t lu[8] = {-1, -1, -1, -1, -1, -1, -1, -1};     // l, u, log_s, p, sign_s, l_mask, u_mask, eye
  int64_t zc_scale = -1;
};

struct DenseW { ConvW conv; int64_t bn_w, bn_b, bn_rm, bn_rv; int64_t scale, shift; int cin; };

struct LevelW {
  int C = 0;                      // flow channels at this level (after squeeze)
  std::vector<StepW> steps;
  ConvW split; int64_t split_gain = -1; int64_t split_scale = -1;
  // encoder side
  ConvW trans; bool has_trans = false;
  std::vector<DenseW> dense;
  int nf_in = 0, nf_out = 0;      // dense block channels in / out
  ConvW cond;
  // hoisted conditioning tables (one LF input shared by all samples): packed weight slices [9][cond][OP]
  int64_t hoist_wd = -1, hoist_wh = -1;
  int hoist_opd = 0, hoist_oph = 0;
  int64_t hoist_od = -1, hoist_oh = -1;            // the same slices as OIHW (phase 1) ...
  struct HoistChunk { ConvW w; int col0; };
  std::vector<HoistChunk> hoist_f16[2];            // ... packed for conv3x3_f16 in phase 2 (per-sample hoisting),
                                                   // in column chunks of at most 256 outputs
};

}  // namespace tmg

using namespace tmg;

struct tmg_model {
  tmg_config cfg;
  int device = 0;
  std::vector<ParamEntry> entries;
  int64_t n_params = 0;
  int64_t n_packed = 0;
  std::vector<LevelW> levels;
  ConvW in_conv, in_conv3, out_conv;
  int Cz = 0;
  int n_steps = 0;
  int cmax = 0;
  std::vector<PackJob> jobs;
  PackJob* jobs_dev = nullptr;
  std::vector<PackJob> jobs2;       // phase 2: sources are in the packed buffer (written by phase 1)
  PackJob* jobs2_dev = nullptr;
  float* packed = nullptr;          // derived weights
  int64_t step_const_off = 0;       // inside packed
  float* params = nullptr;          // borrowed flat parameter buffer
  bool ready = false;
  int precision = TMG_PREC_FP32;    // which kernels run the heavy 3x3 convolutions
  // training tape bookkeeping: which (level, step) of the last training forward recorded its coupling-network
  // intermediates, and the configuration it ran under (a backward under another configuration is refused)
  std::vector<LuTabEntry> lu_tab;   // deferred LU backward: one entry per flow step
  LuTabEntry* lu_tab_dev = nullptr;
  unsigned* sync_dev = nullptr;     // zero-initialised, self-resetting words for single-launch reductions (absmax);
                                    // word 32: sticky fp16-operand overflow flag of the level-resident flow kernel
  LevelStep* lvsteps_dev[TMG_MAX_LEVELS] = {nullptr};   // per level: the plain steps in reverse execution order
  int64_t* lvwmx_dev[TMG_MAX_LEVELS] = {nullptr};       // per level: offsets of the steps' tensor-core mix operands (or null)
  // CUDA graphs of the per-time-step backward (~1 400 launches each): keyed by every pointer / shape the launch sequence
  // depends on; a key is run eagerly the first time it is seen, captured the second time, replayed from then on
  struct BwdGraph { int seen = 0; cudaGraphExec_t exec = nullptr; int64_t kernels = 0; };
  std::map<std::vector<uint64_t>, BwdGraph> bwd_graphs;
  int n_graphs = 0;
  int64_t n_replays = 0, n_eager = 0;
  cudaStream_t gstream = nullptr;   // capture / replay stream (stream capture is not allowed on the legacy default stream)
  cudaEvent_t gev_in = nullptr, gev_out = nullptr;
  // What a training forward recorded, keyed by the TAPE BUFFER it recorded into (a backward may run on an older tape
  // after later forwards under another precision / shape: the per-step emit flags and the configuration belong to the
  // tape, not to the model).  An entry is replaced when its buffer is recorded into again.
  struct TapeInfo { int sig[4] = {0, 0, 0, -1}; std::vector<std::vector<char>> emit; uint64_t serial = 0; };
  std::unordered_map<const void*, TapeInfo> tapes;
  uint64_t tape_serial = 0;
  std::mutex tape_mu;               // the backward runs on autograd's thread
};

namespace tmg {

// ------------------------------------------------------------------ model construction
struct Builder {
  tmg_model& m;
  int64_t add(const std::string& name, std::initializer_list<int64_t> shape) {
    ParamEntry e{name, m.n_params, 1, (int)shape.size(), {1, 1, 1, 1}};
    int k = 0;
    for (int64_t d : shape) { e.dims[k++] = d; e.numel *= d; }
    m.entries.push_back(e);
    m.n_params += e.numel;
    return e.off;
  }
  int64_t pack_alloc(int64_t n) {
    int64_t o = m.n_packed;
    m.n_packed += (n + 3) / 4 * 4;   // keep 16 B alignment
    return o;
  }
  ConvW conv(const std::string& name, int O, int I, bool bias, bool tc = false) {
    ConvW c;
    c.O = O; c.I = I; c.OP = (O + 3) / 4 * 4;
    c.w_param = add(name + ".weight", {O, I, 3, 3});
    if (bias) c.b_param = add(name + ".bias", {O});
    c.w_pack = pack_alloc((int64_t)9 * I * c.OP);
    PackJob j{};
    j.type = JOB_CONVW; j.a = O; j.b = I; j.opad = c.OP;
    for (auto& s : j.src) s = -1;
    j.src[0] = c.w_param; j.dst[0] = c.w_pack;
    m.jobs.push_back(j);
    if (tc) {
      c.NP = tc_npad(O);
      const size_t total = tc_packed_floats(I, c.NP);
      c.w_pack_tc = pack_alloc((int64_t)total);
      const int nparts = (int)((total + 32767) / 32768);
      for (int part = 0; part < nparts; ++part) {
        PackJob t{};
        t.type = JOB_CONVW_TC; t.a = O; t.b = I; t.opad = c.NP; t.part = part; t.nparts = nparts;
        for (auto& s : t.src) s = -1;
        t.src[0] = c.w_param; t.dst[0] = c.w_pack_tc;
        m.jobs.push_back(t);
      }
    }
    return c;
  }
  void conv_f16_job(ConvW& c, int n0, int n1, int n2) {
    const int nch[3] = {n0, n1, n2};
    const int nsrc = n2 > 0 ? 3 : (n1 > 0 ? 2 : 1);
    const int NP = tc_npad(c.O);
    c.NP = NP;
    c.f16_nch[0] = n0; c.f16_nch[1] = n1; c.f16_nch[2] = n2;
    c.w_pack_f16 = pack_alloc((int64_t)convf16_packed_floats(nch, nsrc, NP));
    c.inv_f16 = pack_alloc(1);
    PackJob j{};
    j.type = JOB_CONV_F16; j.a = c.O; j.b = c.I; j.opad = NP; j.nch0 = n0; j.nch1 = n1; j.nd = n2;
    for (auto& s : j.src) s = -1;
    j.src[0] = c.w_param; j.dst[0] = c.w_pack_f16; j.dst[1] = c.inv_f16;
    m.jobs.push_back(j);
  }
  // fp16 packing of `full` over the sources [first n0 input channels | n1 channels after skipping `skip`]
  ConvW conv_f16_skip_job(const ConvW& full, int n0, int skip, int n1) {
    ConvW c = full;
    c.w_pack = c.w_pack_tc = c.w_pack_f16t = -1;
    const int nch[3] = {n0, n1, 0};
    c.NP = tc_npad(c.O);
    c.f16_nch[0] = n0; c.f16_nch[1] = n1; c.f16_nch[2] = 0;
    c.w_pack_f16 = pack_alloc((int64_t)convf16_packed_floats(nch, 2, c.NP));
    c.inv_f16 = pack_alloc(1);
    PackJob j{};
    j.type = JOB_CONV_F16; j.a = c.O; j.b = c.I; j.opad = c.NP; j.nch0 = n0; j.nch1 = n1; j.nd = 0; j.part = skip;
    for (auto& s : j.src) s = -1;
    j.src[0] = c.w_param; j.dst[0] = c.w_pack_f16; j.dst[1] = c.inv_f16;
    m.jobs.push_back(j);
    return c;
  }
  // two-pass gate packing over `nsrc` sources given as (first input channel, channels), in staging order
  ConvW gate2p_job(const ConvW& full, const int* c0, const int* nch, int nsrc) {
    ConvW c = full;
    c.w_pack = c.w_pack_tc = c.w_pack_f16t = -1;
    c.NP = 128;
    int n3[3] = {0, 0, 0};
    for (int i = 0; i < nsrc; ++i) { n3[i] = nch[i]; c.f16_nch[i] = nch[i]; }
    c.w_pack_f16 = pack_alloc((int64_t)lstm_gate_packed_floats(n3, nsrc));
    c.inv_f16 = pack_alloc(1);
    PackJob j{};
    j.type = JOB_GATE2P; j.a = c.O; j.b = c.I; j.opad = 128; j.nch0 = n3[0]; j.nch1 = n3[1]; j.nd = n3[2];
    for (auto& s : j.src) s = -1;
    j.src[0] = c.w_param; j.dst[0] = c.w_pack_f16; j.dst[1] = c.inv_f16;
    for (int i = 0; i < nsrc; ++i) j.src[1 + i] = c0[i];
    m.jobs.push_back(j);
    return c;
  }
  int64_t slice_job(const ConvW& full, int c0, int n, int& op, int pass_rh = 0) {
    op = (full.O + 3) / 4 * 4;
    const int64_t dst = pack_alloc((int64_t)9 * n * op);
    PackJob j{};
    j.type = JOB_SLICE; j.a = full.O; j.b = full.I; j.opad = op; j.nch0 = c0; j.nch1 = n; j.part = pass_rh;
    for (auto& s : j.src) s = -1;
    j.src[0] = full.w_param; j.dst[0] = dst;
    m.jobs.push_back(j);
    return dst;
  }
  void conv_f16t_job(ConvW& c) {             // data-gradient weights: K = O, N = I (each forward source padded to 4 columns)
    if (c.f16_nch[0] + c.f16_nch[1] + c.f16_nch[2] != c.I) return;
    const int NP = tc_npad((c.f16_nch[0] + 3) / 4 * 4 + (c.f16_nch[1] + 3) / 4 * 4 + (c.f16_nch[2] + 3) / 4 * 4);
    if (NP > 256) return;
    const int nch[1] = {c.O};
    c.NPt = NP;
    c.w_pack_f16t = pack_alloc((int64_t)convf16_packed_floats(nch, 1, NP));
    c.inv_f16t = pack_alloc(1);
    PackJob j{};
    j.type = JOB_CONV_F16_T; j.a = c.O; j.b = c.I; j.opad = NP;
    j.nch0 = c.f16_nch[0]; j.nch1 = c.f16_nch[1]; j.nd = c.f16_nch[2];
    for (auto& s : j.src) s = -1;
    j.src[0] = c.w_param; j.dst[0] = c.w_pack_f16t; j.dst[1] = c.inv_f16t;
    m.jobs.push_back(j);
  }
  void coupling_jobs(StepW& st, int nch0, int nch1, int C) {
    const int PT = cpl_planes(nch0, nch1);
    const int NK1 = (PT + 1) / 2 * 2, NPL = (PT + 2) / 2 * 2;
    st.cpl_nch0 = nch0; st.cpl_nch1 = nch1; st.cpl_npad = tc_npad(C);
    st.cpl_w1 = pack_alloc((int64_t)2 * NK1 * 16 * 4);
    st.cpl_w2 = pack_alloc((int64_t)2 * NPL * 16 * 4);
    st.cpl_w3 = pack_alloc((int64_t)9 * 2 * NPL * st.cpl_npad * 4);
    auto job = [&](int type, const ConvW& w, int nplanes, int nd, int64_t dst) {
      PackJob j{};
      j.type = type; j.a = w.O; j.b = w.I; j.opad = st.cpl_npad; j.nplanes = nplanes;
      j.nch0 = nch0; j.nch1 = nch1; j.nd = nd;
      for (auto& s : j.src) s = -1;
      j.src[0] = w.w_param; j.dst[0] = dst;
      m.jobs.push_back(j);
    };
    job(JOB_CPL_W12, st.d1, NK1, 0, st.cpl_w1);
    job(JOB_CPL_W12, st.d2, NPL, 1, st.cpl_w2);
    job(JOB_CPL_W3, st.zc, NPL, 2, st.cpl_w3);
  }
  void step2_jobs(StepW& st, int nch0, int nch1, int C) {
    st.s2_nch0 = nch0; st.s2_nch1 = nch1;
    st.s2_wE = pack_alloc((int64_t)step2_wE_floats(nch0, nch1));
    st.s2_wZ = pack_alloc((int64_t)step2_wZ_floats(nch0, nch1, C));
    st.s2_misc = pack_alloc(16);
    PackJob j{};
    j.type = JOB_STEP2; j.a = C; j.b = st.d1.I; j.opad = tc_npad(C); j.nch0 = nch0; j.nch1 = nch1;
    for (auto& s : j.src) s = -1;
    j.src[0] = st.d1.w_param; j.src[1] = st.d2.w_param; j.src[2] = st.zc.w_param;
    j.dst[0] = st.s2_wE; j.dst[1] = st.s2_wZ; j.dst[2] = st.s2_misc;
    m.jobs.push_back(j);
    if (nch0 + 2 <= 8 && nch1 > 0) {     // plain step of a narrow level: compact operand unit, two taps per MMA
      st.s2c_wE = pack_alloc((int64_t)2 * 2 * 32 * 16 / 4);
      st.s2c_wZ = pack_alloc((int64_t)5 * 2 * 2 * tc_npad(C) * 16 / 4);
      PackJob k = j;
      k.type = JOB_STEP2C;
      k.dst[0] = st.s2c_wE; k.dst[1] = st.s2c_wZ; k.dst[2] = -1;
      m.jobs.push_back(k);
    }
  }
  int64_t gain(int64_t scale_param) {
    int64_t o = pack_alloc(1);
    PackJob j{};
    j.type = JOB_GAIN;
    for (auto& s : j.src) s = -1;
    j.src[0] = scale_param; j.dst[0] = o;
    m.jobs.push_back(j);
    return o;
  }
};

static int build_model(tmg_model& m) {
  const tmg_config& c = m.cfg;
  if (c.n_levels < 1 || c.n_levels > TMG_MAX_LEVELS || c.in_features < 1 || c.out_features < 1 ||
      c.cond_features < 1 || c.cglow_upscale < 1 || c.growth_rate < 1 || c.init_features < 2 ||
      c.rec_features < 1) {
    set_error("bad TMGlow configuration");
    return TMG_ERR_BAD_CONFIG;
  }
  Builder B{m};
  const int L = c.n_levels;
  m.levels.resize(L);
  for (const char* nm : {"in_mu", "in_std", "out_mu", "out_std"}) B.add(nm, {3});   // tmGlow.py:371-374

  // ---- encoder (tmGlow.py:53-102,131-186)
  m.in_conv = B.conv("encoder.first_encoder.In_conv", c.init_features / 2, c.in_features, false);
  m.in_conv3 = B.conv("encoder.first_encoder.In_conv3", c.init_features, c.init_features / 2, false);
  m.Cz = c.out_features << L;   // enc_out_features = out_features * 2^L  (tmGlow.py:348)
  // channel bookkeeping first (out_conv is registered before the blocks in the state_dict)
  {
    int nf = c.init_features;
    for (int i = 0; i < L; ++i) {
      if (i > 0) nf = nf / 2;
      m.levels[i].nf_in = nf;
      nf += c.enc_blocks[i] * c.growth_rate;
      m.levels[i].nf_out = nf;
    }
    m.out_conv = B.conv("encoder.out_conv.0", 2 * m.Cz, nf, false);
  }
  for (int i = 0; i < L; ++i) {
    LevelW& lv = m.levels[i];
    if (c.enc_blocks[i] < 0) { set_error("bad enc_blocks"); return TMG_ERR_BAD_CONFIG; }
    std::string bp = "encoder.encoding_blocks." + std::to_string(i) + ".";
    if (i > 0) {
      lv.has_trans = true;
      lv.trans = B.conv(bp + "encode_conv" + std::to_string(i) + ".conv1", lv.nf_in, m.levels[i - 1].nf_out, false);
    }
    for (int l = 1; l <= c.enc_blocks[i]; ++l) {
      std::string lp = bp + "encode_dense_block" + std::to_string(i) + ".denselayer" + std::to_string(l) + ".";
      DenseW d{};
      d.cin = lv.nf_in + (l - 1) * c.growth_rate;
      d.bn_w = B.add(lp + "norm1.weight", {d.cin});
      d.bn_b = B.add(lp + "norm1.bias", {d.cin});
      d.bn_rm = B.add(lp + "norm1.running_mean", {d.cin});
      d.bn_rv = B.add(lp + "norm1.running_var", {d.cin});
      d.conv = B.conv(lp + "conv1", c.growth_rate, d.cin, false);
      d.scale = B.pack_alloc(d.cin);
      d.shift = B.pack_alloc(d.cin);
      PackJob j{};
      j.type = JOB_BN; j.a = d.cin;
      for (auto& s : j.src) s = -1;
      j.src[0] = d.bn_w; j.src[1] = d.bn_b; j.src[2] = d.bn_rm; j.src[3] = d.bn_rv;
      j.dst[0] = d.scale; j.dst[1] = d.shift;
      m.jobs.push_back(j);
      lv.dense.push_back(d);
    }
  }
  for (int i = 0; i < L; ++i) {
    m.levels[i].cond = B.conv("encoder.cond_convs." + std::to_string(i) + ".0", c.cond_features, m.levels[i].nf_out, false);
    B.conv_f16_job(m.levels[i].cond, m.levels[i].nf_out, 0, 0);      // 60 % of the encoder FLOPs: tensor cores in the f16 modes
  }
  if (tc_npad(m.out_conv.O) <= 256) B.conv_f16_job(m.out_conv, m.out_conv.I, 0, 0);

  // ---- flow blocks (flowLSTMBlock.py:244-278)
  int C = c.out_features;
  m.n_steps = 0;
  for (int b = 0; b < L; ++b) {
    LevelW& lv = m.levels[b];
    C *= 4;
    lv.C = C;
    if (C % 4 != 0 || C > kMaxC) {
      set_error("level %d has %d flow channels; supported: multiples of 4 up to %d", b, C, kMaxC);
      return TMG_ERR_UNSUPPORTED;
    }
    m.cmax = C > m.cmax ? C : m.cmax;
    const int n = c.glow_blocks[b];
    if (n < 1) { set_error("glow_blocks[%d] must be >= 1", b); return TMG_ERR_BAD_CONFIG; }
    const int cin_t = C / 2 + c.cond_features;
    for (int s = 1; s <= n; ++s) {
      StepW st;
      st.kind = (s == n) ? STEP_LSTM : (s == 1 ? STEP_UNNORMED : STEP_PLAIN);
      std::string sp = "glow.flow_blocks." + std::to_string(b) + ".revlayers.affine_layer" + std::to_string(s) + ".";
      if (st.kind != STEP_UNNORMED) {
        st.norm_w = B.add(sp + "norm.weight", {C, 1, 1});
        st.norm_b = B.add(sp + "norm.bias", {C, 1, 1});
      }
      if (st.kind == STEP_LSTM) {           // declared, never used (flowLSTMBlock.py:170)
        B.add(sp + "norm2.weight", {C, 1, 1});
        B.add(sp + "norm2.bias", {C, 1, 1});
      }
      PackJob j{};
      j.type = JOB_1X1; j.a = C;
      j.src[0] = B.add(sp + "conv.l", {C, C});
      j.src[1] = B.add(sp + "conv.u", {C, C});
      j.src[2] = B.add(sp + "conv.log_s", {C});
      j.src[3] = B.add(sp + "conv.p", {C, C});
      j.src[4] = B.add(sp + "conv.sign_s", {C});
      j.src[5] = B.add(sp + "conv.l_mask", {C, C});
      j.src[6] = B.add(sp + "conv.u_mask", {C, C});
      j.src[7] = B.add(sp + "conv.eye", {C, C});
      B.add(sp + "conv.log_s_old", {C});       // cache key of the reference (glowConv.py:209); unused here
      j.src[8] = st.norm_w;
      st.W = B.pack_alloc((int64_t)C * C);
      st.Wi = B.pack_alloc((int64_t)C * C);
      j.dst[3] = -1;
      // tensor-core mix of the level-resident kernel (wide levels; with several tiles per thread the round trip through the MMA
      // issuer of one tile runs under the coupling of the next)
      if (C >= mixmma_min_c() && C % 8 == 0 && !mixmma_off()) { st.Wmx = B.pack_alloc(level_mix_floats(C)); j.dst[3] = st.Wmx; }
      st.const_idx = m.n_steps++;
      for (int q = 0; q < 8; ++q) st.lu[q] = j.src[q];
      j.dst[0] = st.W; j.dst[1] = st.Wi; j.dst[2] = -1;   // step_const offset patched below
      j.b = st.const_idx;
      m.jobs.push_back(j);
      if (st.kind == STEP_LSTM) {
        const int R = c.rec_features;
        st.gate = B.conv(sp + "coupling.resid_lstm.convLSTM.conv", 4 * R, cin_t + R, true, 4 * R <= 256);
        st.outc = B.conv(sp + "coupling.resid_lstm.out_seq.LSTM_out_conv", cin_t, cin_t + R, true, true);
        if (4 * R <= 256) B.conv_f16_job(st.gate, C / 2, c.cond_features, R);
        B.conv_f16_job(st.outc, C / 2, c.cond_features, R);
        st.d1 = B.conv(sp + "coupling.dense_nn.dense_block.denselayer1.conv1", 1, cin_t, false);
        st.d2 = B.conv(sp + "coupling.dense_nn.dense_block.denselayer2.conv1", 1, cin_t + 1, false);
        int64_t sc = B.add(sp + "coupling.out_conv.zero_conv.scale", {1, 1, 1, 1});
        st.zc_scale = sc;
        st.zc = B.conv(sp + "coupling.out_conv.zero_conv.conv", C, cin_t + 2, true, true);
        st.zc_gain = B.gain(sc);
        B.coupling_jobs(st, cin_t, 0, C);
        B.step2_jobs(st, cin_t, 0, C);
        B.conv_f16_job(st.d1, cin_t, 0, 0); B.conv_f16_job(st.d2, cin_t, 1, 0); B.conv_f16_job(st.zc, cin_t, 2, 0);
        B.conv_f16t_job(st.gate); B.conv_f16t_job(st.outc);
        if (4 * R <= 256) {
          st.gate_nc = B.conv_f16_skip_job(st.gate, C / 2, c.cond_features, R);
          if (R == 64 && !gate2p_off()) {
            // staging order of the two-pass kernel: h first (its 8-channel planes pair up into aligned 64-byte reads)
            const int c0a[3] = {cin_t, C / 2, 0}, na[3] = {R, c.cond_features, C / 2};
            const int c0b[2] = {cin_t, 0}, nb[2] = {R, C / 2};
            st.gate2p = B.gate2p_job(st.gate, c0a, na, 3);
            st.gate2p_nc = B.gate2p_job(st.gate, c0b, nb, 2);
            st.gate_hw = st.gate2p_hw = B.slice_job(st.gate, C / 2, c.cond_features, st.gate_hop, R / 2);
          } else {
            st.gate_hw = B.slice_job(st.gate, C / 2, c.cond_features, st.gate_hop);
          }
          st.outc_nc = B.conv_f16_skip_job(st.outc, C / 2, c.cond_features, R);
          st.outc_hw = B.slice_job(st.outc, C / 2, c.cond_features, st.outc_hop);
        }
      } else {
        st.d1 = B.conv(sp + "coupling.coupling_nn.dense_block.denselayer1.conv1", 1, cin_t, false);
        st.d2 = B.conv(sp + "coupling.coupling_nn.dense_block.denselayer2.conv1", 1, cin_t + 1, false);
        int64_t sc = B.add(sp + "coupling.coupling_nn.zero_conv.scale", {1, 1, 1, 1});
        st.zc_scale = sc;
        st.zc = B.conv(sp + "coupling.coupling_nn.zero_conv.conv", C, cin_t + 2, true, true);
        st.zc_gain = B.gain(sc);
        B.coupling_jobs(st, C / 2, c.cond_features, C);
        B.step2_jobs(st, C / 2, c.cond_features, C);
        B.conv_f16_job(st.d1, C / 2, c.cond_features, 0); B.conv_f16_job(st.d2, C / 2, c.cond_features, 1);
        B.conv_f16_job(st.zc, C / 2, c.cond_features, 2);
      }
      B.conv_f16t_job(st.d1); B.conv_f16t_job(st.d2); B.conv_f16t_job(st.zc);
      lv.steps.push_back(st);
    }
    // hoisted conditioning tables: one weight slice per plain step, concatenated along Cout
    lv.hoist_opd = (2 * n + 3) / 4 * 4;
    lv.hoist_oph = (n * C + 3) / 4 * 4;
    lv.hoist_wd = B.pack_alloc((int64_t)9 * c.cond_features * lv.hoist_opd);
    lv.hoist_wh = B.pack_alloc((int64_t)9 * c.cond_features * lv.hoist_oph);
    lv.hoist_od = B.pack_alloc((int64_t)9 * c.cond_features * lv.hoist_opd);
    lv.hoist_oh = B.pack_alloc((int64_t)9 * c.cond_features * lv.hoist_oph);
    for (int kind = 0; kind < 2; ++kind) {        // phase-2 fp16 packing of the concatenated slices (conv3x3_f16.cu)
      const int total = kind ? n * C : 2 * n;       // real columns (the 4-padding columns are never read)
      const int unit = kind ? C : 2;                // keep a step's columns in one chunk
      const int per = std::max(unit, 256 / unit * unit);
      for (int col0 = 0; col0 < total; col0 += per) {
        LevelW::HoistChunk hc{};
        hc.col0 = col0;
        ConvW& cw = hc.w;
        cw.O = std::min(per, total - col0); cw.I = c.cond_features; cw.NP = tc_npad(cw.O);
        const int nch[3] = {c.cond_features, 0, 0};
        cw.w_pack_f16 = B.pack_alloc((int64_t)convf16_packed_floats(nch, 1, cw.NP));
        cw.inv_f16 = B.pack_alloc(1);
        PackJob j{};
        j.type = JOB_CONV_F16; j.a = cw.O; j.b = cw.I; j.opad = cw.NP; j.nch0 = c.cond_features; j.nch1 = 0; j.nd = 0;
        for (auto& q : j.src) q = -1;
        j.src[0] = (kind ? lv.hoist_oh : lv.hoist_od) + (int64_t)col0 * c.cond_features * 9;   // offset into the PACKED buffer
        j.dst[0] = cw.w_pack_f16; j.dst[1] = cw.inv_f16;
        m.jobs2.push_back(j);
        lv.hoist_f16[kind].push_back(hc);
      }
    }
    for (int s = 0; s < n; ++s) {
      const StepW& st = lv.steps[s];
      if (st.kind == STEP_LSTM) continue;       // its coupling net reads the LSTM output, not cond
      for (int kind = 0; kind < 2; ++kind) {
        PackJob j{};
        j.type = JOB_HOIST; j.a = C; j.b = c.cond_features; j.nch0 = C / 2; j.nd = kind; j.part = s; j.nparts = n;
        j.opad = kind ? lv.hoist_oph : lv.hoist_opd;
        for (auto& q : j.src) q = -1;
        j.src[0] = st.d1.w_param; j.src[1] = st.d2.w_param; j.src[2] = st.zc.w_param;
        j.dst[0] = kind ? lv.hoist_wh : lv.hoist_wd;
        j.dst[1] = kind ? lv.hoist_oh : lv.hoist_od;
        m.jobs.push_back(j);
      }
    }
    std::string pp = "glow.flow_blocks." + std::to_string(b) + ".split.latent_encoder.conv2d";
    int64_t sc = B.add(pp + ".scale", {1, 1, 1, 1});
    lv.split_scale = sc;
    lv.split = B.conv(pp + ".conv", C, C / 2, true);
    B.conv_f16_job(lv.split, C / 2, 0, 0);
    B.conv_f16t_job(lv.split);
    lv.split_gain = B.gain(sc);
    C = C / 2;
  }
  m.step_const_off = B.pack_alloc(m.n_steps);
  for (auto& j : m.jobs)
    if (j.type == JOB_1X1) j.dst[2] = m.step_const_off + j.b;
  return TMG_OK;
}

// ------------------------------------------------------------------ workspace plan
struct Plan {
  int B, Bx, h, w, H, W, L;     // Bx: batch of the LF input / encoder (1 when one input is shared by all samples)
  int shared;
  int eh[TMG_MAX_LEVELS], ew[TMG_MAX_LEVELS];   // encoder level sizes
  int Hl[TMG_MAX_LEVELS], Wl[TMG_MAX_LEVELS];   // flow level sizes
  int ctas;                                     // log-det partials per slot
  int nslots;
  size_t total;
  // offsets in floats
  size_t xn, e0, db[TMG_MAX_LEVELS], cc, cond[TMG_MAX_LEVELS], zo_pre, zout;
  size_t bn_mean, bn_var, bn_scale, bn_shift;
  size_t y[TMG_MAX_LEVELS], y2[TMG_MAX_LEVELS], hr, d, gates, u0, ldp, scratch_in, scratch_cond, scratch_out;
  size_t dc_all[TMG_MAX_LEVELS], hc_all[TMG_MAX_LEVELS];
  size_t dcT[TMG_MAX_LEVELS], hcT[TMG_MAX_LEVELS];   // the same tables, plane-transposed for the level-resident kernel
  size_t gcT[TMG_MAX_LEVELS];                         // gate table, plane-transposed + bias (two-pass gate kernel)
  size_t gc[TMG_MAX_LEVELS], oc[TMG_MAX_LEVELS];     // shared LF input: conditioning part of the ConvLSTM gate / output convs
};

static int make_plan(const tmg_model& m, int B, int h, int w, Plan& p, bool shared = false) {
  const tmg_config& c = m.cfg;
  const int L = c.n_levels;
  if (B < 1 || h < 1 || w < 1) { set_error("bad batch/input size"); return TMG_ERR_BAD_SHAPE; }
  p.B = B; p.h = h; p.w = w; p.L = L;
  p.shared = shared ? 1 : 0; p.Bx = shared ? 1 : B;
  p.H = h * c.cglow_upscale; p.W = w * c.cglow_upscale;
  if (p.H % (1 << L) || p.W % (1 << L)) {      // flowUtils.py:112 assert
    set_error("high-fidelity size %dx%d not divisible by 2^%d", p.H, p.W, L);
    return TMG_ERR_BAD_SHAPE;
  }
  int eh = h, ew = w;
  for (int l = 0; l < L; ++l) {
    eh = (eh + 1) / 2; ew = (ew + 1) / 2;       // 3x3 stride-2 pad-1 convolution
    p.eh[l] = eh; p.ew[l] = ew;
    p.Hl[l] = p.H >> (l + 1); p.Wl[l] = p.W >> (l + 1);
    if (eh * c.cglow_upscale != p.Hl[l] || ew * c.cglow_upscale != p.Wl[l]) {
      set_error("level %d: conditioning map %dx%d does not match flow map %dx%d", l,
                eh * c.cglow_upscale, ew * c.cglow_upscale, p.Hl[l], p.Wl[l]);
      return TMG_ERR_BAD_SHAPE;
    }
  }
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += align_up(n, 64); return o; };
  const size_t Bz = (size_t)B, Bx = (size_t)p.Bx;
  p.xn = take(Bx * h * w * c.in_features);
  p.e0 = take(Bx * h * w * (c.init_features / 2));
  size_t cc_max = 0, mx_y = 0, mx_g = 0, mx_u = 0, mx_pix = 0;
  int nf_max = 0;
  for (int l = 0; l < L; ++l) {
    const LevelW& lv = m.levels[l];
    p.db[l] = take(Bx * p.eh[l] * p.ew[l] * lv.nf_out);
    p.cond[l] = take(Bx * p.Hl[l] * p.Wl[l] * c.cond_features);
    p.dc_all[l] = take(Bx * p.Hl[l] * p.Wl[l] * lv.hoist_opd);
    p.hc_all[l] = take(Bx * p.Hl[l] * p.Wl[l] * lv.hoist_oph);
    {
      const StepW& ls = lv.steps.back();
      p.gc[l] = take(shared && ls.gate_hw >= 0 ? (size_t)p.Hl[l] * p.Wl[l] * ls.gate_hop : 16);
      p.oc[l] = take(shared && ls.outc_hw >= 0 ? (size_t)p.Hl[l] * p.Wl[l] * ls.outc_hop : 16);
      p.gcT[l] = take(shared && ls.gate2p_hw >= 0 ? (size_t)p.Hl[l] * p.Wl[l] * ls.gate_hop : 16);
    }
    p.dcT[l] = take(Bx * p.Hl[l] * p.Wl[l] * 2 * lv.steps.size());
    p.hcT[l] = take(Bx * p.Hl[l] * p.Wl[l] * lv.C * lv.steps.size());
    p.y[l] = take(Bz * p.Hl[l] * p.Wl[l] * lv.C);
    p.y2[l] = take(Bz * p.Hl[l] * p.Wl[l] * lv.C);
    cc_max = std::max(cc_max, Bx * p.eh[l] * p.ew[l] * c.cond_features);
    size_t pix = Bz * p.Hl[l] * p.Wl[l];
    mx_pix = std::max(mx_pix, pix);
    mx_y = std::max(mx_y, pix * lv.C);
    mx_g = std::max(mx_g, pix * 4 * c.rec_features);
    mx_u = std::max(mx_u, pix * (size_t)((lv.C / 2 + c.cond_features + 3) / 4 * 4));
    nf_max = std::max(nf_max, lv.nf_out);
  }
  p.cc = take(cc_max);
  p.zo_pre = take(Bx * p.eh[L - 1] * p.ew[L - 1] * 2 * m.Cz);
  p.zout = take(Bx * p.Hl[L - 1] * p.Wl[L - 1] * 2 * m.Cz);
  p.bn_mean = take(nf_max); p.bn_var = take(nf_max);
  p.bn_scale = take(nf_max); p.bn_shift = take(nf_max);
  p.hr = take(mx_y);
  p.d = take(mx_pix * 2);
  p.gates = take(mx_g);
  p.u0 = take(mx_u);
  p.ctas = std::max(cdiv(p.Hl[0] * p.Wl[0], kPixTile), coupling_tc_tiles(p.Hl[0], p.Wl[0]));
  p.ctas = std::max(p.ctas, step2_ld_slots(p.Hl[0], p.Wl[0]));
  p.nslots = 1;
  for (int l = 0; l < L; ++l) p.nslots += c.glow_blocks[l] + 1;
  p.ldp = take(Bz * p.nslots * p.ctas);
  // scratch for the NCHW single-operator entry points
  p.scratch_in = take(mx_y);
  p.scratch_cond = take(mx_pix * c.cond_features);
  p.scratch_out = take(mx_y);
  p.total = off * sizeof(float);
  return TMG_OK;
}

static inline bool prec_tc(int p) { return p != TMG_PREC_FP32; }
// TMG_NO_RESIDENT=1: one launch per flow step (flow_step_f16.cu) even where the level-resident kernel applies (A/B runs)
// TMG_NO_LSTM_HOIST=1: keep the conditioning rows inside the ConvLSTM convolutions (A/B runs)
static inline bool nc_off() {
  static const bool off = [] { const char* e = getenv("TMG_NO_LSTM_HOIST"); return e && e[0] == '1'; }();
  return off;
}
// TMG_TAIL_TC=1: LSTM tails the fused fp16 step kernel does not fit through coupling_tc_kernel as before (A/B runs)
static inline bool tail_tc() {
  static const bool on = [] { const char* e = getenv("TMG_TAIL_TC"); return e && e[0] == '1'; }();
  return on;
}
// TMG_NO_FUSED_UNSQ=1: CheckerSqueeze.reverse as its own launch after the level-resident kernel (A/B runs)
static inline bool unsq_off() {
  static const bool off = [] { const char* e = getenv("TMG_NO_FUSED_UNSQ"); return e && e[0] == '1'; }();
  return off;
}
static inline bool resident_off() {
  static const bool off = [] { const char* e = getenv("TMG_NO_RESIDENT"); return e && e[0] == '1'; }();
  return off;
}
static inline bool prec_split(int p) { return p == TMG_PREC_TF32X3 || p == TMG_PREC_F16X3; }
static inline bool prec_f16(int p) { return p == TMG_PREC_F16X3 || p == TMG_PREC_F16; }

struct Ctx {
  tmg_model& m;
  Plan p;
  float* ws;
  cudaStream_t st;
  bool hoist_ready = false;     // the per-step conditioning tables dc_all / hc_all of this call are filled
  bool unfused = false;         // backward recompute: every conv on its own (the intermediates are needed), no fused epilogues
  bool f16_small = false;   // route the Cout = 1 / zero convs of a step through conv3x3_f16.cu (LSTM tail the fused step kernel does not fit)
  float* emit_d = nullptr;      // training forward: where the fused step kernel records relu(d1), relu(d2) ...
  float* emit_h = nullptr;      // ... and the coupling-network output h of the current step (tape)
  bool emitted = false;         // set by run_step when the launch recorded them
  const float* P() const { return m.params; }
  const float* Q() const { return m.packed; }
};

static int run_conv(Ctx& c, int tag, const ConvW& w, const ConvSrc* srcs, int nsrc, int B, int Hin, int Win, int stride,
                    bool replicate, int act, int64_t gain_pack, const float* bn_scale, const float* bn_shift,
                    float* out, int out_cstride, int out_coff) {
  ConvArgs a{};
  for (int i = 0; i < nsrc; ++i) a.src[i] = srcs[i];
  a.nsrc = nsrc;
  a.bn_scale = bn_scale; a.bn_shift = bn_shift;
  a.w = c.Q() + w.w_pack; a.cin_w = w.I; a.cout_w = w.OP;
  a.bias = w.b_param >= 0 ? c.P() + w.b_param : nullptr;
  a.gain = gain_pack >= 0 ? c.Q() + gain_pack : nullptr;
  a.act = act;
  a.out = out; a.out_cstride = out_cstride; a.out_coff = out_coff; a.cout = w.O;
  a.B = B; a.Hin = Hin; a.Win = Win; a.stride = stride;
  a.Hout = stride == 1 ? Hin : (Hin + 1) / 2;
  a.Wout = stride == 1 ? Win : (Win + 1) / 2;
  a.pad_replicate = replicate ? 1 : 0;
  if ((c.unfused || c.f16_small) && prec_f16(c.m.precision) && w.w_pack_f16 >= 0 && stride == 1 && !bn_scale) {
    bool match = true;
    for (int i = 0; i < 3; ++i) match = match && (i < nsrc ? srcs[i].nch : 0) == w.f16_nch[i];
    ConvF16Args t{};
    for (int i = 0; i < nsrc; ++i) t.src[i] = srcs[i];
    t.nsrc = nsrc;
    t.wpk = c.Q() + w.w_pack_f16; t.inv_scale = c.Q() + w.inv_f16; t.npad = w.NP;
    t.bias = a.bias; t.gain = a.gain; t.act = act;
    t.out = out; t.out_cstride = out_cstride; t.out_coff = out_coff; t.cout = w.O;
    t.B = B; t.H = Hin; t.W = Win; t.pad_replicate = a.pad_replicate; t.x3 = prec_split(c.m.precision) ? 1 : 0; t.overflow = c.m.sync_dev ? c.m.sync_dev + 32 : nullptr;
    if (match && convf16_supported(t)) {
      const double M = (double)B * Hin * Win;
      ProfScope ps(c.st, tag, 2.0 * M * w.O * 9.0 * w.I, 4.0 * ((double)B * Hin * Win * w.I + M * w.O));
      return launch_conv3x3_f16(t, c.st);
    }
  }
  if (c.m.precision != TMG_PREC_FP32 && w.w_pack_tc >= 0 && stride == 1 && !bn_scale && Win + 2 <= 512 && !c.unfused) {
    TcConvArgs t{};
    for (int i = 0; i < nsrc; ++i) t.src[i] = srcs[i];
    t.nsrc = nsrc; t.cin = w.I; t.wpk = c.Q() + w.w_pack_tc; t.npad = w.NP;
    t.bias = a.bias; t.gain = a.gain; t.act = act;
    t.out = out; t.out_cstride = out_cstride; t.out_coff = out_coff; t.cout = w.O;
    t.B = B; t.H = Hin; t.W = Win; t.pad_replicate = a.pad_replicate;
    t.split3 = prec_split(c.m.precision) ? 1 : 0;
    const double M = (double)B * Hin * Win;
    ProfScope ps(c.st, tag, 2.0 * M * w.O * 9.0 * w.I, 4.0 * ((double)B * Hin * Win * w.I + M * w.O));
    return launch_conv3x3_tc(t, c.st);
  }
  // algorithmic work: 2*M*N*K flops; bytes = read every input channel once + write the outputs
  const double M = (double)B * a.Hout * a.Wout;
  ProfScope ps(c.st, tag, 2.0 * M * w.O * 9.0 * w.I, 4.0 * ((double)B * Hin * Win * w.I + M * w.O));
  return launch_conv3x3(a, c.st);
}

// ------------------------------------------------------------------ encoder
static int run_encoder(Ctx& c, const float* x, bool bn_train, float bn_momentum = 0.1f) {
  const tmg_config& g = c.m.cfg;
  const Plan& p = c.p;
  float* ws = c.ws;
  const int B = p.Bx, L = p.L;
  PermArgs pa{};
  pa.src = x; pa.dst = ws + p.xn; pa.mode = PERM_NCHW_TO_NHWC;
  pa.B = B; pa.C = g.in_features; pa.H = p.h; pa.W = p.w; pa.dst_cstride = g.in_features;
  TMG_TRY(launch_permute(pa, c.st));
  ConvSrc s0{ws + p.xn, g.in_features, 0, g.in_features, 0};
  TMG_TRY(run_conv(c, PROF_CONV_ENC, c.m.in_conv, &s0, 1, B, p.h, p.w, 1, false, 0, -1, nullptr, nullptr,
                   ws + p.e0, g.init_features / 2, 0));
  ConvSrc s1{ws + p.e0, g.init_features / 2, 0, g.init_features / 2, 1};
  TMG_TRY(run_conv(c, PROF_CONV_ENC, c.m.in_conv3, &s1, 1, B, p.h, p.w, 2, false, 0, -1, nullptr, nullptr,
                   ws + p.db[0], c.m.levels[0].nf_out, 0));
  for (int i = 0; i < L; ++i) {
    const LevelW& lv = c.m.levels[i];
    float* db = ws + p.db[i];
    const int eh = p.eh[i], ew = p.ew[i];
    if (i > 0) {
      const LevelW& pv = c.m.levels[i - 1];
      ConvSrc st{ws + p.db[i - 1], pv.nf_out, 0, pv.nf_out, 1};
      TMG_TRY(run_conv(c, PROF_CONV_ENC, lv.trans, &st, 1, B, p.eh[i - 1], p.ew[i - 1], 2, false, 0, -1, nullptr, nullptr,
                       db, lv.nf_out, 0));
    }
    int stats_done = 0;
    for (size_t l = 0; l < lv.dense.size(); ++l) {
      const DenseW& d = lv.dense[l];
      const float* sc; const float* sh;
      if (bn_train) {
        BnStatArgs sa{};
        sa.x = db; sa.cstride = lv.nf_out; sa.c0 = stats_done; sa.n = d.cin - stats_done;
        sa.N = (int64_t)B * eh * ew; sa.mean = ws + p.bn_mean; sa.var = ws + p.bn_var;
        TMG_TRY(launch_bn_stats(sa, c.st));
        stats_done = d.cin;
        BnFoldArgs fa{};
        fa.mean = ws + p.bn_mean; fa.var = ws + p.bn_var;
        fa.w = c.P() + d.bn_w; fa.b = c.P() + d.bn_b;
        fa.run_mean = c.m.params + d.bn_rm; fa.run_var = c.m.params + d.bn_rv;
        fa.scale = ws + p.bn_scale; fa.shift = ws + p.bn_shift;
        fa.n = d.cin; fa.N = sa.N; fa.eps = 1e-5f; fa.momentum = bn_momentum;
        TMG_TRY(launch_bn_fold_train(fa, c.st));
        sc = ws + p.bn_scale; sh = ws + p.bn_shift;
      } else {
        sc = c.Q() + d.scale; sh = c.Q() + d.shift;
      }
      ConvSrc sd{db, lv.nf_out, 0, d.cin, 1};
      TMG_TRY(run_conv(c, PROF_CONV_ENC, d.conv, &sd, 1, B, eh, ew, 1, false, 0, -1, sc, sh, db, lv.nf_out, d.cin));
    }
    ConvSrc sc{db, lv.nf_out, 0, lv.nf_out, 0};
    const bool up = g.cglow_upscale > 1;
    auto enc_f16 = [&](const ConvW& w, float* out, int cstride) -> int {      // 1: done, 0: not applicable, < 0: error
      if (!prec_f16(c.m.precision) || w.w_pack_f16 < 0) return 0;
      ConvF16Args t{};
      t.src[0] = sc; t.nsrc = 1;
      t.wpk = c.Q() + w.w_pack_f16; t.inv_scale = c.Q() + w.inv_f16; t.npad = w.NP; t.cout = w.O;
      t.out = out; t.out_cstride = cstride; t.out_coff = 0;
      t.B = B; t.H = eh; t.W = ew; t.x3 = prec_split(c.m.precision) ? 1 : 0; t.overflow = c.m.sync_dev ? c.m.sync_dev + 32 : nullptr;
      if (!convf16_supported(t)) return 0;
      const double M = (double)B * eh * ew;
      ProfScope ps(c.st, PROF_CONV_ENC, 2.0 * M * w.O * 9.0 * w.I, 4.0 * M * (w.I + w.O));
      const int rc = launch_conv3x3_f16(t, c.st);
      return rc == TMG_OK ? 1 : rc;
    };
    {
      const int r = enc_f16(lv.cond, up ? ws + p.cc : ws + p.cond[i], g.cond_features);
      if (r < 0) return r;
      if (r == 0)
        TMG_TRY(run_conv(c, PROF_CONV_ENC, lv.cond, &sc, 1, B, eh, ew, 1, false, 0, -1, nullptr, nullptr,
                         up ? ws + p.cc : ws + p.cond[i], g.cond_features, 0));
    }
    if (up) {
      UpsampleArgs ua{ws + p.cc, ws + p.cond[i], B, eh, ew, g.cond_features, g.cglow_upscale};
      TMG_TRY(launch_upsample(ua, c.st));
    }
    if (i == L - 1) {
      const int r = enc_f16(c.m.out_conv, up ? ws + p.zo_pre : ws + p.zout, 2 * c.m.Cz);
      if (r < 0) return r;
      if (r == 0)
        TMG_TRY(run_conv(c, PROF_CONV_ENC, c.m.out_conv, &sc, 1, B, eh, ew, 1, false, 0, -1, nullptr, nullptr,
                         up ? ws + p.zo_pre : ws + p.zout, 2 * c.m.Cz, 0));
      if (up) {
        UpsampleArgs ua{ws + p.zo_pre, ws + p.zout, B, eh, ew, 2 * c.m.Cz, g.cglow_upscale};
        TMG_TRY(launch_upsample(ua, c.st));
      }
    }
  }
  return TMG_OK;
}

// ------------------------------------------------------------------ coupling network of one step -> HR
static int run_coupling_nn(Ctx& c, int level, const StepW& s, int B, int Hl, int Wl, const float* Y,
                           const float* cond, const float* h_in, const float* c_in, float* h_out, float* c_out,
                           bool nn_only_lstm = false) {
  const tmg_config& g = c.m.cfg;
  const Plan& p = c.p;
  float* ws = c.ws;
  const int C = c.m.levels[level].C, cf = g.cond_features, R = g.rec_features;
  const int cin_t = C / 2 + cf;
  float* D = ws + p.d;
  float* HR = ws + p.hr;
  ConvSrc src[3];
  int nt;   // sources that make up "t"
  if (s.kind == STEP_LSTM) {
    if (!h_out || !c_out) { set_error("LSTM step needs h_out/c_out buffers"); return TMG_ERR_NULL; }
    const int sh = c.p.shared;
    ConvSrc gs[3] = {{Y, C, 0, C / 2, 0}, {cond, cf, 0, cf, 0, sh}, {h_in, R, 0, R, 0}};
    bool gate_done = false, out_done = false;
    const int u0s_f = (cin_t + 3) / 4 * 4;
    // one LF input shared by all samples: the conditioning rows leave the two convolutions (K shrinks by the 32 conditioning
    // channels: 7 -> 5 K-steps at level 0) and come back as a per-pixel addend evaluated once per call (run_hoist)
    const bool nc = sh && c.hoist_ready && s.gate_nc.w_pack_f16 >= 0 && s.gate_hw >= 0 && !c.unfused && !nc_off();
    if (prec_f16(c.m.precision) && s.gate2p.w_pack_f16 >= 0 && !c.unfused) {
      // two-pass kernel: the cell-update epilogue of one half of the gate columns overlaps the MMAs of the other half
      ConvF16Args t{};
      const bool nc2 = nc && s.gate2p_hw >= 0;
      t.src[0] = ConvSrc{h_in, R, 0, R, 0};             // zero states: null -> staged as zeros
      if (nc2) {
        t.src[1] = gs[0]; t.nsrc = 2;
        t.wpk = c.Q() + s.gate2p_nc.w_pack_f16; t.inv_scale = c.Q() + s.gate2p_nc.inv_f16;
        t.addend = ws + p.gcT[level]; t.addend_stride = 0;      // bias folded into the table
      } else {
        t.src[1] = gs[1]; t.src[2] = gs[0]; t.nsrc = 3;
        t.wpk = c.Q() + s.gate2p.w_pack_f16; t.inv_scale = c.Q() + s.gate2p.inv_f16;
        t.bias = c.P() + s.gate.b_param;
      }
      t.npad = 256; t.cout = s.gate.O;
      t.B = B; t.H = Hl; t.W = Wl; t.x3 = prec_split(c.m.precision) ? 1 : 0;
      t.lstm_R = R; t.c_prev = c_in; t.h_out = h_out; t.c_out = c_out;
      if (lstm_gate_f16_supported(t)) {
        const double M = (double)B * Hl * Wl;
        ProfScope ps(c.st, PROF_CONV_GATE, 2.0 * M * s.gate.O * 9.0 * s.gate.I, 4.0 * (M * s.gate.I + M * R * (c_in ? 3.0 : 2.0)));
        TMG_TRY(launch_lstm_gate_f16(t, c.st));
        gate_done = true;
      }
    }
    if (!gate_done && prec_f16(c.m.precision) && s.gate.w_pack_f16 >= 0 && !c.unfused) {
      const bool nc = sh && c.hoist_ready && s.gate_nc.w_pack_f16 >= 0 && s.gate_hw >= 0 && s.gate2p_hw < 0 && !c.unfused && !nc_off();
      ConvF16Args t{};
      const int ns = h_in ? 3 : 2;
      for (int i = 0; i < 3; ++i) t.src[i] = gs[i];
      if (!h_in) t.src[2].p = nullptr;        // zero states: the plane is staged as zeros
      t.nsrc = 3; (void)ns;
      t.wpk = c.Q() + s.gate.w_pack_f16; t.inv_scale = c.Q() + s.gate.inv_f16; t.npad = s.gate.NP;
      if (nc) {
        t.src[1] = t.src[2]; t.src[2] = ConvSrc{}; t.nsrc = 2;
        t.wpk = c.Q() + s.gate_nc.w_pack_f16; t.inv_scale = c.Q() + s.gate_nc.inv_f16;
        t.addend = ws + p.gc[level]; t.addend_stride = s.gate_hop;
      }
      t.bias = c.P() + s.gate.b_param; t.cout = s.gate.O;
      t.B = B; t.H = Hl; t.W = Wl; t.x3 = prec_split(c.m.precision) ? 1 : 0; t.overflow = c.m.sync_dev ? c.m.sync_dev + 32 : nullptr;
      t.lstm_R = R; t.c_prev = c_in; t.h_out = h_out; t.c_out = c_out;
      if (convf16_supported(t)) {
        const double M = (double)B * Hl * Wl;
        ProfScope ps(c.st, PROF_CONV_GATE, 2.0 * M * s.gate.O * 9.0 * s.gate.I, 4.0 * (M * s.gate.I + M * R * (c_in ? 3.0 : 2.0)));
        TMG_TRY(launch_conv3x3_f16(t, c.st));
        gate_done = true;
      }
    }
    if (gate_done) {
      ConvF16Args t{};
      ConvSrc os2[3] = {{Y, C, 0, C / 2, 0}, {cond, cf, 0, cf, 0, sh}, {h_out, R, 0, R, 0}};
      for (int i = 0; i < 3; ++i) t.src[i] = os2[i];
      t.nsrc = 3;
      t.wpk = c.Q() + s.outc.w_pack_f16; t.inv_scale = c.Q() + s.outc.inv_f16; t.npad = s.outc.NP;
      if (nc) {
        t.src[1] = t.src[2]; t.src[2] = ConvSrc{}; t.nsrc = 2;
        t.wpk = c.Q() + s.outc_nc.w_pack_f16; t.inv_scale = c.Q() + s.outc_nc.inv_f16;
        t.addend = ws + p.oc[level]; t.addend_stride = s.outc_hop;
      }
      t.bias = c.P() + s.outc.b_param; t.cout = s.outc.O; t.act = 1;
      t.out = ws + p.u0; t.out_cstride = u0s_f; t.out_coff = 0;
      t.B = B; t.H = Hl; t.W = Wl; t.x3 = prec_split(c.m.precision) ? 1 : 0; t.overflow = c.m.sync_dev ? c.m.sync_dev + 32 : nullptr;
      if (s.outc.w_pack_f16 >= 0 && convf16_supported(t)) {
        const double M = (double)B * Hl * Wl;
        ProfScope ps(c.st, PROF_CONV_OUT, 2.0 * M * s.outc.O * 9.0 * s.outc.I, 4.0 * (M * s.outc.I + M * s.outc.O));
        TMG_TRY(launch_conv3x3_f16(t, c.st));
        out_done = true;
      }
    }
    if (gate_done) {
      // fall through to the output conv below if it was not done
    } else if (c.m.precision != TMG_PREC_FP32 && s.gate.w_pack_tc >= 0 && R % 16 == 0 && Wl + 2 <= 512 && !c.unfused) {
      // gate conv on tcgen05 with the ConvLSTM cell update fused into its epilogue (gates never touch HBM)
      TcConvArgs t{};
      const int ns = h_in ? 3 : 2;
      for (int i = 0; i < ns; ++i) t.src[i] = gs[i];
      t.nsrc = ns; t.cin = s.gate.I; t.wpk = c.Q() + s.gate.w_pack_tc; t.npad = s.gate.NP;
      t.bias = c.P() + s.gate.b_param; t.cout = s.gate.O;
      t.B = B; t.H = Hl; t.W = Wl; t.split3 = prec_split(c.m.precision) ? 1 : 0;
      t.lstm_R = R; t.c_prev = c_in; t.h_out = h_out; t.c_out = c_out;
      const double M = (double)B * Hl * Wl;
      ProfScope ps(c.st, PROF_CONV_GATE, 2.0 * M * s.gate.O * 9.0 * s.gate.I,
                   4.0 * (M * s.gate.I + M * R * (c_in ? 3.0 : 2.0)));
      TMG_TRY(launch_conv3x3_tc(t, c.st));
    } else {
      TMG_TRY(run_conv(c, PROF_CONV_GATE, s.gate, gs, h_in ? 3 : 2, B, Hl, Wl, 1, false, 0, -1, nullptr, nullptr,
                       ws + p.gates, 4 * R, 0));
      LstmArgs la{ws + p.gates, c_in, h_out, c_out, (int64_t)B * Hl * Wl * R, R};
      ProfScope ps(c.st, PROF_LSTM_PW, 20.0 * la.n, 4.0 * la.n * (c_in ? 7.0 : 6.0));
      TMG_TRY(launch_lstm_pointwise(la, c.st));
    }
    ConvSrc os[3] = {{Y, C, 0, C / 2, 0}, {cond, cf, 0, cf, 0, sh}, {h_out, R, 0, R, 0}};
    const int u0s = (cin_t + 3) / 4 * 4;       // padded pixel stride: 16-byte aligned planes for the fused kernel
    if (!out_done)
      TMG_TRY(run_conv(c, PROF_CONV_OUT, s.outc, os, 3, B, Hl, Wl, 1, false, 1, -1, nullptr, nullptr, ws + p.u0, u0s, 0));
    if (nn_only_lstm) return TMG_OK;
    src[0] = ConvSrc{ws + p.u0, u0s, 0, cin_t, 1};
    nt = 1;
  } else {
    if (nn_only_lstm) return TMG_OK;
    src[0] = ConvSrc{Y, C, 0, C / 2, 1};
    src[1] = ConvSrc{cond, cf, 0, cf, 1, c.p.shared};
    nt = 2;
  }
  TMG_TRY(run_conv(c, PROF_CONV_DENSE1, s.d1, src, nt, B, Hl, Wl, 1, false, 0, -1, nullptr, nullptr, D, 2, 0));
  src[nt] = ConvSrc{D, 2, 0, 1, 1};
  TMG_TRY(run_conv(c, PROF_CONV_DENSE1, s.d2, src, nt + 1, B, Hl, Wl, 1, false, 0, -1, nullptr, nullptr, D, 2, 1));
  src[nt] = ConvSrc{D, 2, 0, 2, 1};
  TMG_TRY(run_conv(c, PROF_CONV_ZERO, s.zc, src, nt + 1, B, Hl, Wl, 1, true, 0, s.zc_gain, nullptr, nullptr, HR, C, 0));
  return TMG_OK;
}

static int run_pointwise(Ctx& c, int level, float* Y, bool coupling, const StepW* mix, bool reverse,
                         int B, int HW, float* ld_slot) {
  PointArgs a{};
  a.y = Y;
  a.hr = coupling ? c.ws + c.p.hr : nullptr;
  if (mix) {
    a.wmat = c.Q() + (reverse ? mix->W : mix->Wi);
    if (mix->kind != STEP_UNNORMED) { a.nw = c.P() + mix->norm_w; a.nb = c.P() + mix->norm_b; }
  }
  a.reverse = reverse ? 1 : 0;
  a.B = B; a.HW = HW; a.C = c.m.levels[level].C;
  a.ld_part = ld_slot; a.ld_stride = c.p.nslots * c.p.ctas;
  // algorithmic bytes: read + write the flow state, read the coupling-net output
  const double px = (double)B * HW;
  ProfScope ps(c.st, PROF_POINTWISE, px * a.C * (mix ? 2.0 * a.C : 0.0) + px * a.C * 6.0,
               4.0 * px * a.C * (coupling ? 3.0 : 2.0));
  return launch_flow_pointwise(a, c.st);
}

// One flow step on the current flow state *Y (alternate buffer *Y2).
//   reverse: coupling_rev(step) -> W -> ActNorm_rev            (mix = this step)
//   forward: coupling_fwd(step) -> ActNorm_fwd -> W^-1 of `mix` (= the NEXT step, or null)
// Tensor-core precisions run the fused kernel (one launch, result in *Y2, pointers swapped);
// fp32 runs the CUDA-core kernels in place.
static int run_step(Ctx& c, int level, const StepW& st, const StepW* mix, bool reverse, int B, int Hl, int Wl,
                    float*& Y, float*& Y2, const float* cond, const float* h_in, const float* c_in, float* h_out,
                    float* c_out, float* ld_slot) {
  const tmg_config& g = c.m.cfg;
  const int C = c.m.levels[level].C, cf = g.cond_features, HW = Hl * Wl;
  if (prec_f16(c.m.precision) && st.s2_wE >= 0) {
    Step2Args a{};
    if (st.kind == STEP_LSTM) {
      const int cin_t = C / 2 + cf;
      a.src[0] = ConvSrc{c.ws + c.p.u0, (cin_t + 3) / 4 * 4, 0, cin_t, 1};
      a.nsrc = 1;
    } else {
      a.src[0] = ConvSrc{Y, C, 0, C / 2, 1};
      a.src[1] = ConvSrc{cond, cf, 0, cf, 1, c.p.shared};
      a.nsrc = 2;
      if (c.hoist_ready) {
        const LevelW& lv = c.m.levels[level];
        const int sidx = (int)(&st - lv.steps.data());
        a.hoist = 1;
        a.dc = c.ws + c.p.dc_all[level] + 2 * sidx; a.dc_stride = lv.hoist_opd;
        a.hc = c.ws + c.p.hc_all[level] + (size_t)sidx * C; a.hc_stride = lv.hoist_oph;
        a.hoist_bstride = c.p.shared ? 0 : 1;          // tables per sample (distinct LF inputs) or shared by all
      }
    }
    a.wE = c.Q() + st.s2_wE; a.wZ = c.Q() + st.s2_wZ; a.wmisc = c.Q() + st.s2_misc;
    a.bias3 = c.P() + st.zc.b_param; a.gain3 = c.Q() + st.zc_gain;
    a.C = C; a.y_in = Y; a.y_out = Y2;
    if (mix) {
      a.wmat = c.Q() + (reverse ? mix->W : mix->Wi);
      if (mix->kind != STEP_UNNORMED) { a.nw = c.P() + mix->norm_w; a.nb = c.P() + mix->norm_b; }
    }
    a.reverse = reverse ? 1 : 0;
    a.ld_part = ld_slot; a.ld_stride = c.p.nslots * c.p.ctas;
    a.B = B; a.H = Hl; a.W = Wl; a.x3 = prec_split(c.m.precision) ? 1 : 0;
    a.overflow = c.m.sync_dev ? c.m.sync_dev + 32 : nullptr;
    if (st.kind != STEP_LSTM && c.emit_d && c.emit_h) { a.d_emit = c.emit_d; a.h_emit = c.emit_h; }
    if (step2_supported(a)) {
      if (st.kind == STEP_LSTM) TMG_TRY(run_coupling_nn(c, level, st, B, Hl, Wl, Y, cond, h_in, c_in, h_out, c_out, true));
      const double px = (double)B * HW;
      const int cin_t = C / 2 + cf;
      ProfScope ps(c.st, PROF_STEP_FUSED, 2.0 * px * 9.0 * (cin_t + (cin_t + 1) + (double)C * (cin_t + 2)) + 2.0 * px * C * C,
                   4.0 * px * (2.0 * C + cf));
      int rc = launch_step2(a, c.st);
      if (rc == TMG_OK) { float* t = Y; Y = Y2; Y2 = t; c.emitted = a.d_emit != nullptr; }
      return rc;
    }
  }
  if (prec_f16(c.m.precision) && st.kind == STEP_LSTM && st.d1.w_pack_f16 >= 0 && st.zc.w_pack_f16 >= 0 && !c.unfused && !tail_tc()) {
    // LSTM tail the fused fp16 step kernel has no room for (C = 48: 4 K-steps of u0 + all 9 taps of the Z weights exceed the
    // shared memory of an SM): the three convolutions of the coupling net through conv3x3_f16.cu + the pointwise step, instead
    // of the first-generation TF32 kernel (one CTA per sample: 1.5 ms at S = 4096, level 2)
    c.f16_small = true;
    int rc = run_coupling_nn(c, level, st, B, Hl, Wl, Y, cond, h_in, c_in, h_out, c_out);
    c.f16_small = false;
    TMG_TRY(rc);
    return run_pointwise(c, level, Y, true, mix, reverse, B, HW, ld_slot);
  }
  if (c.m.precision != TMG_PREC_FP32 && st.cpl_w3 >= 0) {
    CouplingArgs a{};
    if (st.kind == STEP_LSTM) {
      TMG_TRY(run_coupling_nn(c, level, st, B, Hl, Wl, Y, cond, h_in, c_in, h_out, c_out, true));
      const int cin_t = C / 2 + cf;
      a.src[0] = ConvSrc{c.ws + c.p.u0, (cin_t + 3) / 4 * 4, 0, cin_t, 1};
      a.nsrc = 1;
    } else {
      a.src[0] = ConvSrc{Y, C, 0, C / 2, 1};
      a.src[1] = ConvSrc{cond, cf, 0, cf, 1, c.p.shared};
      a.nsrc = 2;
    }
    a.w1 = c.Q() + st.cpl_w1; a.w2 = c.Q() + st.cpl_w2; a.w3 = c.Q() + st.cpl_w3; a.npad = st.cpl_npad;
    a.bias3 = c.P() + st.zc.b_param; a.gain3 = c.Q() + st.zc_gain;
    a.C = C; a.y_in = Y; a.y_out = Y2;
    if (mix) {
      a.wmat = c.Q() + (reverse ? mix->W : mix->Wi);
      if (mix->kind != STEP_UNNORMED) { a.nw = c.P() + mix->norm_w; a.nb = c.P() + mix->norm_b; }
    }
    a.reverse = reverse ? 1 : 0;
    a.ld_part = ld_slot; a.ld_stride = c.p.nslots * c.p.ctas;
    a.B = B; a.H = Hl; a.W = Wl; a.split3 = prec_split(c.m.precision) ? 1 : 0;
    const double px = (double)B * HW;
    const int cin_t = C / 2 + cf;
    // algorithmic work of the whole step: three convs + 1x1; bytes: read x + cond, write y
    ProfScope ps(c.st, PROF_STEP_FUSED, 2.0 * px * 9.0 * (cin_t + (cin_t + 1) + (double)C * (cin_t + 2)) + 2.0 * px * C * C,
                 4.0 * px * (2.0 * C + cf));
    int rc = launch_coupling_tc(a, c.st);
    if (rc == TMG_OK) { float* t = Y; Y = Y2; Y2 = t; }
    return rc;
  }
  TMG_TRY(run_coupling_nn(c, level, st, B, Hl, Wl, Y, cond, h_in, c_in, h_out, c_out));
  return run_pointwise(c, level, Y, true, mix, reverse, B, HW, ld_slot);
}

static int run_split_prior(Ctx& c, int level, int B, int Hl, int Wl, const float* Y) {
  const LevelW& lv = c.m.levels[level];
  ConvSrc s{Y, lv.C, 0, lv.C / 2, 0};
  if (prec_f16(c.m.precision) && lv.split.w_pack_f16 >= 0) {
    ConvF16Args t{};
    t.src[0] = s; t.nsrc = 1;
    t.wpk = c.Q() + lv.split.w_pack_f16; t.inv_scale = c.Q() + lv.split.inv_f16; t.npad = lv.split.NP;
    t.bias = c.P() + lv.split.b_param; t.gain = c.Q() + lv.split_gain; t.act = 2; t.cout = lv.split.O;
    t.out = c.ws + c.p.hr; t.out_cstride = lv.C; t.out_coff = 0;
    t.B = B; t.H = Hl; t.W = Wl; t.pad_replicate = 1; t.x3 = prec_split(c.m.precision) ? 1 : 0; t.overflow = c.m.sync_dev ? c.m.sync_dev + 32 : nullptr;
    if (convf16_supported(t)) {
      const double M = (double)B * Hl * Wl;
      ProfScope ps(c.st, PROF_CONV_SPLIT, 2.0 * M * lv.split.O * 9.0 * lv.split.I, 4.0 * (M * lv.split.I + M * lv.split.O));
      return launch_conv3x3_f16(t, c.st);
    }
  }
  return run_conv(c, PROF_CONV_SPLIT, lv.split, &s, 1, B, Hl, Wl, 1, true, 2, lv.split_gain, nullptr, nullptr,
                  c.ws + c.p.hr, lv.C, 0);
}

// One LF input shared by every sample: the conditioning map's contribution to every plain step's coupling net
// is the same for all samples.  Two exact-fp32 convolutions per level fill, for all steps at once,
//   dc_all[pix][2s], [2s+1] = conv3x3_zero-pad(relu(cond), w1_s / w2_s cond rows)
//   hc_all[pix][s*C + n]    = conv3x3_replicate(relu(cond), zero_conv_s cond rows)       (before bias and gain)
static int run_hoist(Ctx& c) {
  const tmg_config& g = c.m.cfg;
  if (!c.p.shared) {
    // distinct LF inputs: the same tables per sample, as two tensor-core convolutions per level (N = all steps at
    // once) -- the conditioning K-steps leave the 15 step kernels of the level and run as one well-shaped GEMM
    for (int l = 0; l < c.p.L; ++l) {
      const LevelW& lv = c.m.levels[l];
      for (int kind = 0; kind < 2; ++kind) {
        for (const auto& hc : lv.hoist_f16[kind]) {
          const ConvW& cw = hc.w;
          ConvF16Args t{};
          t.src[0] = ConvSrc{c.ws + c.p.cond[l], g.cond_features, 0, g.cond_features, 1, 0};
          t.nsrc = 1;
          t.wpk = c.Q() + cw.w_pack_f16; t.inv_scale = c.Q() + cw.inv_f16; t.npad = cw.NP; t.cout = cw.O;
          t.out = c.ws + (kind ? c.p.hc_all[l] : c.p.dc_all[l]);
          t.out_cstride = kind ? lv.hoist_oph : lv.hoist_opd; t.out_coff = hc.col0;
          t.B = c.p.B; t.H = c.p.Hl[l]; t.W = c.p.Wl[l]; t.pad_replicate = kind;
          t.x3 = prec_split(c.m.precision) ? 1 : 0; t.overflow = c.m.sync_dev ? c.m.sync_dev + 32 : nullptr;
          if (!convf16_supported(t)) { set_error("hoisted conditioning conv: unsupported shape"); return TMG_ERR_UNSUPPORTED; }
          const double M = (double)t.B * t.H * t.W;
          ProfScope ps(c.st, PROF_MISC, 2.0 * M * cw.O * 9.0 * cw.I, 4.0 * M * (cw.I + cw.O));
          TMG_TRY(launch_conv3x3_f16(t, c.st));
        }
      }
    }
    c.hoist_ready = true;
    return TMG_OK;
  }
  for (int l = 0; l < c.p.L; ++l) {
    const LevelW& lv = c.m.levels[l];
    for (int kind = 0; kind < 2; ++kind) {
      ConvArgs a{};
      a.src[0] = ConvSrc{c.ws + c.p.cond[l], g.cond_features, 0, g.cond_features, 1, 1};
      a.nsrc = 1;
      a.w = c.Q() + (kind ? lv.hoist_wh : lv.hoist_wd);
      a.cin_w = g.cond_features; a.cout_w = kind ? lv.hoist_oph : lv.hoist_opd;
      a.out = c.ws + (kind ? c.p.hc_all[l] : c.p.dc_all[l]);
      a.out_cstride = a.cout_w; a.out_coff = 0; a.cout = a.cout_w;
      a.B = 1; a.Hin = c.p.Hl[l]; a.Win = c.p.Wl[l]; a.Hout = a.Hin; a.Wout = a.Win; a.stride = 1;
      a.pad_replicate = kind;
      const double M = (double)a.Hin * a.Win;
      ProfScope ps(c.st, PROF_MISC, 2.0 * M * a.cout * 9.0 * a.cin_w, 4.0 * M * (a.cin_w + a.cout));
      TMG_TRY(launch_conv3x3(a, c.st));
    }
    // conditioning rows of the ConvLSTM gate and output convolutions (zero padding, no ReLU: convLSTM.py:44,129)
    const StepW& ls = lv.steps.back();
    if (ls.kind == STEP_LSTM && ls.gate_hw >= 0) {
      for (int kind = 0; kind < 2; ++kind) {
        ConvArgs a{};
        a.src[0] = ConvSrc{c.ws + c.p.cond[l], g.cond_features, 0, g.cond_features, 0, 1};
        a.nsrc = 1;
        a.w = c.Q() + (kind ? ls.outc_hw : ls.gate_hw);
        a.cin_w = g.cond_features; a.cout_w = kind ? ls.outc_hop : ls.gate_hop;
        a.out = c.ws + (kind ? c.p.oc[l] : c.p.gc[l]);
        a.out_cstride = a.cout_w; a.out_coff = 0; a.cout = a.cout_w;
        a.B = 1; a.Hin = c.p.Hl[l]; a.Win = c.p.Wl[l]; a.Hout = a.Hin; a.Wout = a.Win; a.stride = 1;
        const double M = (double)a.Hin * a.Win;
        ProfScope ps(c.st, PROF_MISC, 2.0 * M * a.cout * 9.0 * a.cin_w, 4.0 * M * (a.cin_w + a.cout));
        TMG_TRY(launch_conv3x3(a, c.st));
      }
      if (ls.gate2p_hw >= 0)      // pass-ordered columns -> [column group of 4][pixel] float4, bias folded in
        TMG_TRY(launch_gate_addend_transpose(c.ws + c.p.gc[l], c.P() + ls.gate.b_param, c.ws + c.p.gcT[l], c.p.Hl[l] * c.p.Wl[l],
                                             ls.gate_hop, g.rec_features, c.st));
    }
  }
  c.hoist_ready = true;
  return TMG_OK;
}

static int finish_logdet(Ctx& c, float* out) {
  LogdetArgs a{};
  a.ld_part = c.ws + c.p.ldp; a.ld_stride = c.p.nslots * c.p.ctas;
  a.step_const = c.Q() + c.m.step_const_off;
  a.n_levels = c.p.L;
  int k = 0;
  for (int l = 0; l < c.p.L; ++l) {
    a.step_begin[l] = k;
    k += (int)c.m.levels[l].steps.size();
    a.hw[l] = c.p.Hl[l] * c.p.Wl[l];
  }
  a.step_begin[c.p.L] = k;
  a.out = out; a.B = c.p.B;
  return launch_logdet_reduce(a, c.st);
}

static int check_common(tmg_model* m, const void* ws, size_t ws_bytes, const Plan& p) {
  if (!m) { set_error("null model"); return TMG_ERR_NULL; }
  if (!m->ready) { set_error("tmg_model_refresh() has not been called"); return TMG_ERR_NOT_READY; }
  if (!ws) { set_error("null workspace"); return TMG_ERR_NULL; }
  if (ws_bytes < p.total) { set_error("workspace too small: %zu < %zu bytes", ws_bytes, p.total); return TMG_ERR_WORKSPACE; }
  if (((uintptr_t)ws & 255) != 0) { set_error("workspace must be 256-byte aligned"); return TMG_ERR_BAD_SHAPE; }
  return TMG_OK;
}

}  // namespace tmg

// =================================================================== C ABI
extern "C" {

int tmg_version(void) { return 110; }
const char* tmg_last_error(void) { return tmg::g_err; }

int tmg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

static const char* kProfNames[PROF_NTAGS] = {"conv_lstm_gates", "conv_lstm_out", "conv_zero", "conv_dense_cout1",
                                              "conv_split_prior", "conv_encoder", "flow_pointwise", "lstm_pointwise",
                                              "gaussian", "permute", "misc", "flow_step_fused", "flow_level_resident"};

int tmg_profile_enable(int on) {
  for (auto& r : tmg::g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  tmg::g_prof.clear();
  tmg::g_prof_on = on != 0;
  return TMG_OK;
}
int tmg_profile_classes(void) { return PROF_NTAGS; }
const char* tmg_profile_class_name(int tag) { return (tag >= 0 && tag < PROF_NTAGS) ? kProfNames[tag] : nullptr; }
int tmg_profile_query(int tag, double* ms, int64_t* launches, double* flops, double* bytes) {
  if (!ms || !launches || !flops || !bytes) { set_error("null argument"); return TMG_ERR_NULL; }
  *ms = 0; *launches = 0; *flops = 0; *bytes = 0;
  for (auto& r : tmg::g_prof) {
    if (r.tag != tag) continue;
    TMG_CUDA_OK(cudaEventSynchronize(r.b));
    float t = 0.f;
    TMG_CUDA_OK(cudaEventElapsedTime(&t, r.a, r.b));
    *ms += t; *launches += 1; *flops += r.flops; *bytes += r.bytes;
  }
  return TMG_OK;
}

int64_t tmg_launch_count(int reset) {
  int64_t v = reset ? tmg::g_launches.exchange(0) : tmg::g_launches.load();
  return v;
}

int tmg_model_create(const tmg_config* cfg, tmg_model** out) {
  if (!cfg || !out) { set_error("null argument"); return TMG_ERR_NULL; }
  tmg_model* m = new tmg_model();
  m->cfg = *cfg;
  int s = build_model(*m);
  if (s != TMG_OK) { delete m; return s; }
  *out = m;
  return TMG_OK;
}

void tmg_model_destroy(tmg_model* m) {
  if (!m) return;
  if (m->jobs_dev) cudaFree(m->jobs_dev);
  if (m->jobs2_dev) cudaFree(m->jobs2_dev);
  if (m->lu_tab_dev) cudaFree(m->lu_tab_dev);
  if (m->sync_dev) cudaFree(m->sync_dev);
  for (auto* p_ : m->lvsteps_dev) if (p_) cudaFree(p_);
  for (auto* p_ : m->lvwmx_dev) if (p_) cudaFree(p_);
  for (auto& kv : m->bwd_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  if (m->gstream) cudaStreamDestroy(m->gstream);
  if (m->gev_in) cudaEventDestroy(m->gev_in);
  if (m->gev_out) cudaEventDestroy(m->gev_out);
  if (m->packed) cudaFree(m->packed);
  delete m;
}

int tmg_model_set_precision(tmg_model* m, int mode) {
  if (!m) { set_error("null model"); return TMG_ERR_NULL; }
  if (mode < TMG_PREC_FP32 || mode > TMG_PREC_F16) {
    set_error("unknown precision mode %d", mode); return TMG_ERR_BAD_CONFIG;
  }
  m->precision = mode;
  return TMG_OK;
}
int tmg_model_get_precision(const tmg_model* m) { return m ? m->precision : -1; }

// Sticky overflow flag of the fp16-operand kernels (modes f16x3 / f16): the fp32 activations are clamped to +-6e4 when they
// are split into fp16 hi/lo halves; a clamp that actually changed a value -- outside the reference's semantics, it means
// ill-conditioned weights -- raises the flag.  Synchronises the stream.  Returns 1 (and sets tmg_last_error) when raised.
int tmg_model_overflow(tmg_model* m, int clear, void* stream) {
  if (!m) { set_error("null model"); return TMG_ERR_NULL; }
  if (!m->sync_dev) return 0;
  unsigned v = 0;
  cudaStream_t st = (cudaStream_t)stream;
  TMG_CUDA_OK(cudaMemcpyAsync(&v, m->sync_dev + 32, sizeof(v), cudaMemcpyDeviceToHost, st));
  TMG_CUDA_OK(cudaStreamSynchronize(st));
  if (v && clear) TMG_CUDA_OK(cudaMemsetAsync(m->sync_dev + 32, 0, sizeof(unsigned), st));
  if (v) set_error("an activation exceeded the fp16 operand range (|v| > 6e4) and was clamped: results are outside the "
                   "reference's fp32 semantics (ill-conditioned weights?); use precision \"fp32\"");
  return v ? 1 : 0;
}

int64_t tmg_model_param_entries(const tmg_model* m) { return m ? (int64_t)m->entries.size() : 0; }
const char* tmg_model_param_name(const tmg_model* m, int64_t i) {
  return (m && i >= 0 && i < (int64_t)m->entries.size()) ? m->entries[i].name.c_str() : nullptr;
}
int64_t tmg_model_param_offset(const tmg_model* m, int64_t i) {
  return (m && i >= 0 && i < (int64_t)m->entries.size()) ? m->entries[i].off : -1;
}
int64_t tmg_model_param_numel(const tmg_model* m, int64_t i) {
  return (m && i >= 0 && i < (int64_t)m->entries.size()) ? m->entries[i].numel : -1;
}
int tmg_model_param_shape(const tmg_model* m, int64_t i, int64_t dims[4]) {
  if (!m || !dims || i < 0 || i >= (int64_t)m->entries.size()) return -1;
  for (int k = 0; k < 4; ++k) dims[k] = m->entries[i].dims[k];
  return m->entries[i].ndim;
}
int64_t tmg_model_param_total(const tmg_model* m) { return m ? m->n_params : 0; }

int tmg_model_refresh(tmg_model* m, float* params, void* stream) {
  if (!m || !params) { set_error("null argument"); return TMG_ERR_NULL; }
  cudaStream_t st = (cudaStream_t)stream;
  if (!m->packed) {     // one-time allocation of the library-owned derived-weight cache
    TMG_CUDA_OK(cudaGetDevice(&m->device));
    TMG_CUDA_OK(cudaMalloc(&m->packed, (size_t)m->n_packed * sizeof(float)));
    TMG_CUDA_OK(cudaMemset(m->packed, 0, (size_t)m->n_packed * sizeof(float)));
    TMG_CUDA_OK(cudaMalloc(&m->jobs_dev, m->jobs.size() * sizeof(PackJob)));
    TMG_CUDA_OK(cudaMemcpy(m->jobs_dev, m->jobs.data(), m->jobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice));
    {
      m->lu_tab.clear();
      for (const LevelW& lv : m->levels)
        for (const StepW& sw : lv.steps) {
          LuTabEntry t{};
          t.C = lv.C;
          for (int q = 0; q < 8; ++q) t.off[q] = sw.lu[q];
          t.norm_w = sw.kind != STEP_UNNORMED ? sw.norm_w : -1;
          m->lu_tab.push_back(t);
        }
      TMG_CUDA_OK(cudaMalloc(&m->sync_dev, 64 * sizeof(unsigned)));
      TMG_CUDA_OK(cudaMemset(m->sync_dev, 0, 64 * sizeof(unsigned)));
      TMG_CUDA_OK(cudaMalloc(&m->lu_tab_dev, m->lu_tab.size() * sizeof(LuTabEntry)));
      TMG_CUDA_OK(cudaMemcpy(m->lu_tab_dev, m->lu_tab.data(), m->lu_tab.size() * sizeof(LuTabEntry), cudaMemcpyHostToDevice));
    }
    for (int l = 0; l < m->cfg.n_levels; ++l) {        // step tables of the level-resident flow kernel (offsets only: static)
      const LevelW& lv = m->levels[l];
      std::vector<LevelStep> tab;
      std::vector<int64_t> wmx;
      for (int s = (int)lv.steps.size() - 2; s >= 0; --s) {
        const StepW& sw = lv.steps[s];
        LevelStep e{};
        e.wE = sw.s2_wE; e.wZ = sw.s2_wZ; e.misc = sw.s2_misc; e.gain = sw.zc_gain; e.W = sw.W;
        e.wEc = sw.s2c_wE; e.wZc = sw.s2c_wZ;
        e.bias = sw.zc.b_param;
        e.nw = sw.kind != STEP_UNNORMED ? sw.norm_w : -1; e.nb = sw.kind != STEP_UNNORMED ? sw.norm_b : -1;
        e.dc_off = s; e.hc_off = s;
        tab.push_back(e);
        wmx.push_back(sw.Wmx);
      }
      if (!tab.empty()) {
        TMG_CUDA_OK(cudaMalloc(&m->lvsteps_dev[l], tab.size() * sizeof(LevelStep)));
        TMG_CUDA_OK(cudaMemcpy(m->lvsteps_dev[l], tab.data(), tab.size() * sizeof(LevelStep), cudaMemcpyHostToDevice));
        if (wmx[0] >= 0) {
          TMG_CUDA_OK(cudaMalloc(&m->lvwmx_dev[l], wmx.size() * sizeof(int64_t)));
          TMG_CUDA_OK(cudaMemcpy(m->lvwmx_dev[l], wmx.data(), wmx.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
        }
      }
    }
    if (!m->jobs2.empty()) {
      TMG_CUDA_OK(cudaMalloc(&m->jobs2_dev, m->jobs2.size() * sizeof(PackJob)));
      TMG_CUDA_OK(cudaMemcpy(m->jobs2_dev, m->jobs2.data(), m->jobs2.size() * sizeof(PackJob), cudaMemcpyHostToDevice));
    }
  }
  m->params = params;
  TMG_TRY(launch_pack(m->jobs_dev, (int)m->jobs.size(), params, m->packed, m->cmax, st));
  if (!m->jobs2.empty()) TMG_TRY(launch_pack(m->jobs2_dev, (int)m->jobs2.size(), m->packed, m->packed, m->cmax, st));
  m->ready = true;
  return TMG_OK;
}

int tmg_model_get_conv1x1(const tmg_model* m, int level, int step, int inverse, float* dst, void* stream) {
  if (!m || !dst) { set_error("null argument"); return TMG_ERR_NULL; }
  if (!m->ready) { set_error("model not refreshed"); return TMG_ERR_NOT_READY; }
  if (level < 0 || level >= m->cfg.n_levels || step < 1 || step > (int)m->levels[level].steps.size()) {
    set_error("bad level/step"); return TMG_ERR_BAD_SHAPE;
  }
  const StepW& s = m->levels[level].steps[step - 1];
  int C = m->levels[level].C;
  TMG_CUDA_OK(cudaMemcpyAsync(dst, m->packed + (inverse ? s.Wi : s.W), (size_t)C * C * sizeof(float),
                              cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return TMG_OK;
}

size_t tmg_workspace_bytes(const tmg_model* m, int B, int h, int w) {
  if (!m) return 0;
  Plan p, q;
  if (make_plan(*m, B, h, w, p) != TMG_OK) return 0;
  if (make_plan(*m, B, h, w, q, true) != TMG_OK) return 0;     // TMG_FLAG_SHARED_X: smaller encoder buffers, extra hoisted tables
  return std::max(p.total, q.total);
}

int tmg_encoder_forward(tmg_model* m, int B, int h, int w, const float* x, float* const* c_out, float* z_out,
                        void* workspace, size_t workspace_bytes, uint32_t flags, void* stream) {
  if (!m) { set_error("null model"); return TMG_ERR_NULL; }
  Plan p;
  TMG_TRY(make_plan(*m, B, h, w, p));
  TMG_TRY(check_common(m, workspace, workspace_bytes, p));
  if (!x) { set_error("null input"); return TMG_ERR_NULL; }
  Ctx c{*m, p, (float*)workspace, (cudaStream_t)stream};
  TMG_TRY(run_encoder(c, x, flags & TMG_FLAG_BN_TRAIN));
  for (int l = 0; l < p.L && c_out; ++l) {
    if (!c_out[l]) continue;
    PermArgs pa{};
    pa.src = c.ws + p.cond[l]; pa.dst = c_out[l]; pa.mode = PERM_NHWC_TO_NCHW;
    pa.B = B; pa.C = m->cfg.cond_features; pa.H = p.Hl[l]; pa.W = p.Wl[l]; pa.src_cstride = pa.C;
    TMG_TRY(launch_permute(pa, c.st));
  }
  if (z_out) {
    PermArgs pa{};
    pa.src = c.ws + p.zout; pa.dst = z_out; pa.mode = PERM_NHWC_TO_NCHW;
    pa.B = B; pa.C = 2 * m->Cz; pa.H = p.Hl[p.L - 1]; pa.W = p.Wl[p.L - 1]; pa.src_cstride = pa.C;
    TMG_TRY(launch_permute(pa, c.st));
  }
  return TMG_OK;
}

// Tape of the training forward, levels 0..L-1, steps 0..n-1, per step (NHWC): the step input Y [px,C], then the
// coupling-network intermediates the fused step kernel records: relu(d1), relu(d2) [px,2] and h [px,C].
static size_t tape_off(const tmg_model& m, const Plan& p, int l, int s) {
  size_t off = 0;
  for (int q = 0; q < l; ++q) off += (size_t)m.levels[q].steps.size() * p.B * p.Hl[q] * p.Wl[q] * (2 * m.levels[q].C + 2);
  return off + (size_t)s * p.B * p.Hl[l] * p.Wl[l] * (2 * m.levels[l].C + 2);
}
static size_t tape_off_d(const tmg_model& m, const Plan& p, int l, int s) {
  return tape_off(m, p, l, s) + (size_t)p.B * p.Hl[l] * p.Wl[l] * m.levels[l].C;
}
static size_t tape_off_h(const tmg_model& m, const Plan& p, int l, int s) {
  return tape_off_d(m, p, l, s) + (size_t)p.B * p.Hl[l] * p.Wl[l] * 2;
}
static size_t tape_floats(const tmg_model& m, const Plan& p) { return tape_off(m, p, p.L, 0); }

static int reconstruct_impl(tmg_model* m, int B, int h, int w, const float* x, const float* const* h_in,
                            const float* const* c_in, const float* const* eps, float* y, float* log_det,
                            float* const* h_out, float* const* c_out, void* workspace, size_t workspace_bytes,
                            uint32_t flags, void* stream, float* tape);

int tmg_reconstruct(tmg_model* m, int B, int h, int w, const float* x, const float* const* h_in,
                    const float* const* c_in, const float* const* eps, float* y, float* log_det,
                    float* const* h_out, float* const* c_out, void* workspace, size_t workspace_bytes,
                    uint32_t flags, void* stream) {
  return reconstruct_impl(m, B, h, w, x, h_in, c_in, eps, y, log_det, h_out, c_out, workspace, workspace_bytes, flags, stream, nullptr);
}

size_t tmg_tape_bytes(const tmg_model* m, int B, int h, int w) {
  if (!m) return 0;
  Plan p;
  if (make_plan(*m, B, h, w, p) != TMG_OK) return 0;
  return tape_floats(*m, p) * sizeof(float);
}

int tmg_reconstruct_train(tmg_model* m, int B, int h, int w, const float* x, const float* const* h_in,
                          const float* const* c_in, const float* const* eps, float* y, float* log_det,
                          float* const* h_out, float* const* c_out, void* tape, size_t tape_bytes, void* workspace,
                          size_t workspace_bytes, uint32_t flags, void* stream) {
  if (!tape || tape_bytes < tmg_tape_bytes(m, B, h, w)) { set_error("tape missing or too small"); return TMG_ERR_WORKSPACE; }
  if (flags & TMG_FLAG_SHARED_X) { set_error("TMG_FLAG_SHARED_X is an inference option"); return TMG_ERR_BAD_CONFIG; }
  return reconstruct_impl(m, B, h, w, x, h_in, c_in, eps, y, log_det, h_out, c_out, workspace, workspace_bytes, flags, stream, (float*)tape);
}

static int reconstruct_impl(tmg_model* m, int B, int h, int w, const float* x, const float* const* h_in,
                            const float* const* c_in, const float* const* eps, float* y, float* log_det,
                            float* const* h_out, float* const* c_out, void* workspace, size_t workspace_bytes,
                            uint32_t flags, void* stream, float* tape) {
  if (!m) { set_error("null model"); return TMG_ERR_NULL; }
  Plan p;
  const bool shared = (flags & TMG_FLAG_SHARED_X) != 0;
  if (shared && (flags & TMG_FLAG_BN_TRAIN)) { set_error("TMG_FLAG_SHARED_X cannot be combined with train-mode BatchNorm"); return TMG_ERR_BAD_CONFIG; }
  TMG_TRY(make_plan(*m, B, h, w, p, shared));
  TMG_TRY(check_common(m, workspace, workspace_bytes, p));
  if (!x || !eps || !y || !log_det || !h_out || !c_out) { set_error("null argument"); return TMG_ERR_NULL; }
  const int L = p.L;
  for (int l = 0; l <= L; ++l) if (!eps[l]) { set_error("eps[%d] is null", l); return TMG_ERR_NULL; }
  Ctx c{*m, p, (float*)workspace, (cudaStream_t)stream};
  float* ws = c.ws;
  const int ldstride = p.nslots * p.ctas;
  TMG_CUDA_OK(cudaMemsetAsync(ws + p.ldp, 0, (size_t)B * ldstride * sizeof(float), c.st));
  TMG_TRY(run_encoder(c, x, flags & TMG_FLAG_BN_TRAIN));
  if (prec_f16(m->precision)) TMG_TRY(run_hoist(c));
  tmg_model::TapeInfo tinfo;
  if (tape) {
    tinfo.emit.resize(L);
    for (int l = 0; l < L; ++l) tinfo.emit[l].assign(m->levels[l].steps.size(), 0);
    tinfo.sig[0] = B; tinfo.sig[1] = h; tinfo.sig[2] = w; tinfo.sig[3] = m->precision;
    std::lock_guard<std::mutex> lk(m->tape_mu);
    m->tapes.erase(tape);            // until this forward is fully issued the buffer holds no valid record
  }

  // top latent: z = cmean + exp(clamp(clog_std)) * eps[L]   (tmGlow.py:460-463); no log-prob term
  {
    const LevelW& lv = m->levels[L - 1];
    GaussArgs ga{};
    ga.prm = ws + p.zout; ga.prm_cstride = 2 * m->Cz; ga.prm_bshared = p.shared;
    ga.val = ws + p.y[L - 1]; ga.val_cstride = lv.C; ga.val_coff = 0;
    ga.eps_in = eps[L]; ga.reverse = 1;
    ga.B = B; ga.HW = p.Hl[L - 1] * p.Wl[L - 1]; ga.n = m->Cz;
    ga.ld_part = nullptr; ga.ld_stride = ldstride;
    TMG_TRY(launch_gaussian(ga, c.st));
  }
  int slot = 0;
  for (int l = L - 1; l >= 0; --l) {
    const LevelW& lv = m->levels[l];
    const int Hl = p.Hl[l], Wl = p.Wl[l], HW = Hl * Wl;
    float* Y = ws + p.y[l];
    float* Y2 = ws + p.y2[l];
    // Split.reverse (flowUtils.py:316-335)
    TMG_TRY(run_split_prior(c, l, B, Hl, Wl, Y));
    GaussArgs ga{};
    ga.prm = ws + p.hr; ga.prm_cstride = lv.C;
    ga.val = Y; ga.val_cstride = lv.C; ga.val_coff = lv.C / 2;
    ga.eps_in = eps[l]; ga.reverse = 1; ga.B = B; ga.HW = HW; ga.n = lv.C / 2;
    ga.ld_part = ws + p.ldp + (size_t)(slot++) * p.ctas; ga.ld_stride = ldstride;
    TMG_TRY(launch_gaussian(ga, c.st));
    // steps n..1 reversed (flowLSTMBlock.py:348-359)
    bool unsq_done = false;
    for (int s = (int)lv.steps.size() - 1; s >= 0; --s) {
      const StepW& st = lv.steps[s];
      if (!tape && s == (int)lv.steps.size() - 2 && s >= 0 && prec_f16(m->precision) && c.hoist_ready && m->lvsteps_dev[l] &&
          !resident_off()) {
        // all plain steps of the level in one launch with the state resident on the SM (flow_level_f16.cu)
        LevelArgs la{};
        la.steps = m->lvsteps_dev[l]; la.nsteps = s + 1;
        la.params = c.P(); la.packed = c.Q();
        la.y_in = Y; la.y_out = Y;
        la.dc = ws + p.dcT[l]; la.hc = ws + p.hcT[l]; la.nsteps_tab = (int)lv.steps.size();
        la.hoist_bstride = p.shared ? 0 : 1;
        la.ld_part = ws + p.ldp + (size_t)slot * p.ctas; la.ld_stride = ldstride;
        la.B = B; la.H = Hl; la.W = Wl; la.C = lv.C; la.x3 = prec_split(m->precision) ? 1 : 0;
        la.nch1 = m->cfg.cond_features;
        la.compact = lv.steps[0].s2c_wE >= 0 ? 1 : 0;
        const int64_t* wmx = m->lvwmx_dev[l];
        la.overflow = m->sync_dev + 32;
        if (level_resident_supported(la, wmx != nullptr)) {
          TMG_TRY(launch_hoist_transpose(ws + p.dc_all[l], lv.hoist_opd, ws + p.hc_all[l], lv.hoist_oph, ws + p.dcT[l], ws + p.hcT[l],
                                         p.Bx, HW, (int)lv.steps.size(), lv.C, c.st));
          const double px = (double)B * HW;
          const int cin_t = lv.C / 2 + m->cfg.cond_features;
          ProfScope ps(c.st, PROF_LEVEL_RES,
                       la.nsteps * (2.0 * px * 9.0 * (cin_t + (cin_t + 1) + (double)lv.C * (cin_t + 2)) + 2.0 * px * lv.C * lv.C),
                       la.nsteps * 4.0 * px * (2.0 * lv.C + m->cfg.cond_features));
          // CheckerSqueeze.reverse of wide levels inside the kernel's final store (the level's result feeds nothing else)
          float* unsq = (l > 0 && lv.C != 12 && lv.C % 8 == 0 && !unsq_off()) ? ws + p.y[l - 1] : nullptr;
          TMG_TRY(launch_level_resident(la, wmx, unsq, unsq ? m->levels[l - 1].C : 0, c.st));
          unsq_done = unsq != nullptr;
          slot += la.nsteps;
          break;
        }
      }
      if (tape) {
        TMG_CUDA_OK(cudaMemcpyAsync(tape + tape_off(*m, p, l, s), Y, (size_t)B * HW * lv.C * sizeof(float), cudaMemcpyDeviceToDevice, c.st));
        c.emit_d = tape + tape_off_d(*m, p, l, s); c.emit_h = tape + tape_off_h(*m, p, l, s); c.emitted = false;
      }
      TMG_TRY(run_step(c, l, st, &st, true, B, Hl, Wl, Y, Y2, ws + p.cond[l], h_in ? h_in[l] : nullptr,
                       c_in ? c_in[l] : nullptr, h_out[l], c_out[l], ws + p.ldp + (size_t)(slot++) * p.ctas));
      if (tape) {
        tinfo.emit[l][s] = c.emitted ? 1 : 0;
        c.emit_d = c.emit_h = nullptr;
      }
    }
    // CheckerSqueeze.reverse (flowUtils.py:124-145)
    PermArgs pa{};
    pa.src = Y; pa.src_cstride = lv.C; pa.src_coff = 0;
    pa.B = B; pa.C = lv.C / 4; pa.H = 2 * Hl; pa.W = 2 * Wl;
    if (l > 0) {
      pa.mode = PERM_UNSQUEEZE_NHWC_TO_NHWC;
      pa.dst = ws + p.y[l - 1]; pa.dst_cstride = m->levels[l - 1].C; pa.dst_coff = 0;
    } else {
      pa.mode = PERM_UNSQUEEZE_NHWC_TO_NCHW;
      pa.dst = y;
    }
    if (!unsq_done) TMG_TRY(launch_permute(pa, c.st));
  }
  if (tape) {
    std::lock_guard<std::mutex> lk(m->tape_mu);
    if (m->tapes.size() >= 512) {      // forget the oldest records (their backward, if any, recomputes the intermediates)
      uint64_t cut = m->tape_serial > 256 ? m->tape_serial - 256 : 0;
      for (auto it = m->tapes.begin(); it != m->tapes.end();) it = it->second.serial <= cut ? m->tapes.erase(it) : std::next(it);
    }
    tinfo.serial = ++m->tape_serial;
    m->tapes[tape] = std::move(tinfo);
  }
  return finish_logdet(c, log_det);
}

int tmg_forward(tmg_model* m, int B, int h, int w, const float* x, const float* y, const float* const* h_in,
                const float* const* c_in, float* z, float* logp, float* const* h_out, float* const* c_out,
                float* const* eps_out, void* workspace, size_t workspace_bytes, uint32_t flags, void* stream) {
  if (!m) { set_error("null model"); return TMG_ERR_NULL; }
  Plan p;
  const bool shared = (flags & TMG_FLAG_SHARED_X) != 0;
  if (shared && (flags & TMG_FLAG_BN_TRAIN)) { set_error("TMG_FLAG_SHARED_X cannot be combined with train-mode BatchNorm"); return TMG_ERR_BAD_CONFIG; }
  TMG_TRY(make_plan(*m, B, h, w, p, shared));
  TMG_TRY(check_common(m, workspace, workspace_bytes, p));
  if (!x || !y || !z || !logp || !h_out || !c_out) { set_error("null argument"); return TMG_ERR_NULL; }
  const int L = p.L;
  Ctx c{*m, p, (float*)workspace, (cudaStream_t)stream};
  float* ws = c.ws;
  const int ldstride = p.nslots * p.ctas;
  TMG_CUDA_OK(cudaMemsetAsync(ws + p.ldp, 0, (size_t)B * ldstride * sizeof(float), c.st));
  TMG_TRY(run_encoder(c, x, flags & TMG_FLAG_BN_TRAIN));
  if (prec_f16(m->precision)) TMG_TRY(run_hoist(c));
  int slot = 0;
  float* yfinal[TMG_MAX_LEVELS] = {nullptr};
  for (int l = 0; l < L; ++l) {
    const LevelW& lv = m->levels[l];
    const int Hl = p.Hl[l], Wl = p.Wl[l], HW = Hl * Wl;
    float* Y = ws + p.y[l];
    float* Y2 = ws + p.y2[l];
    // CheckerSqueeze.forward (flowUtils.py:99-121)
    PermArgs pa{};
    pa.dst = Y; pa.dst_cstride = lv.C; pa.dst_coff = 0;
    pa.B = B; pa.C = lv.C / 4; pa.H = 2 * Hl; pa.W = 2 * Wl;
    if (l == 0) { pa.mode = PERM_SQUEEZE_NCHW_TO_NHWC; pa.src = y; }
    else { pa.mode = PERM_SQUEEZE_NHWC_TO_NHWC; pa.src = yfinal[l - 1]; pa.src_cstride = m->levels[l - 1].C; pa.src_coff = 0; }
    TMG_TRY(launch_permute(pa, c.st));
    // steps 1..n: [ActNorm ->] W^-1 -> coupling; the coupling of step s is fused with the
    // ActNorm/W^-1 of step s+1 (flowLSTMBlock.py:296-308)
    const int n = (int)lv.steps.size();
    TMG_TRY(run_pointwise(c, l, Y, false, &lv.steps[0], false, B, HW, nullptr));
    for (int s = 0; s < n; ++s) {
      const StepW& st = lv.steps[s];
      TMG_TRY(run_step(c, l, st, s + 1 < n ? &lv.steps[s + 1] : nullptr, false, B, Hl, Wl, Y, Y2, ws + p.cond[l],
                       h_in ? h_in[l] : nullptr, c_in ? c_in[l] : nullptr, h_out[l], c_out[l],
                       ws + p.ldp + (size_t)(slot++) * p.ctas));
    }
    // Split.forward (flowUtils.py:292-314)
    TMG_TRY(run_split_prior(c, l, B, Hl, Wl, Y));
    GaussArgs ga{};
    ga.prm = ws + p.hr; ga.prm_cstride = lv.C;
    ga.val = Y; ga.val_cstride = lv.C; ga.val_coff = lv.C / 2;
    ga.eps_out = eps_out ? eps_out[l] : nullptr; ga.reverse = 0; ga.B = B; ga.HW = HW; ga.n = lv.C / 2;
    ga.ld_part = ws + p.ldp + (size_t)(slot++) * p.ctas; ga.ld_stride = ldstride;
    TMG_TRY(launch_gaussian(ga, c.st));
    yfinal[l] = Y;
  }
  // top prior log p(z | cmean, clog_std) and eps0 (tmGlow.py:399-412)
  {
    const LevelW& lv = m->levels[L - 1];
    GaussArgs ga{};
    ga.prm = ws + p.zout; ga.prm_cstride = 2 * m->Cz; ga.prm_bshared = p.shared;
    ga.val = yfinal[L - 1]; ga.val_cstride = lv.C; ga.val_coff = 0;
    ga.eps_out = eps_out ? eps_out[L] : nullptr; ga.val_nchw = z; ga.reverse = 0;
    ga.B = B; ga.HW = p.Hl[L - 1] * p.Wl[L - 1]; ga.n = m->Cz;
    ga.ld_part = ws + p.ldp + (size_t)(slot++) * p.ctas; ga.ld_stride = ldstride;
    TMG_TRY(launch_gaussian(ga, c.st));
  }
  return finish_logdet(c, logp);
}

// ------------------------------------------------------------------ single operators
int tmg_squeeze_forward(const float* x, float* y, int B, int C, int H, int W, void* stream) {
  if (!x || !y) { set_error("null argument"); return TMG_ERR_NULL; }
  if (H % 2 || W % 2) { set_error("squeeze: %dx%d not divisible by 2", H, W); return TMG_ERR_BAD_SHAPE; }   // flowUtils.py:112
  PermArgs pa{};
  pa.src = x; pa.dst = y; pa.mode = PERM_SQUEEZE_NCHW_TO_NCHW; pa.B = B; pa.C = C; pa.H = H; pa.W = W;
  return launch_permute(pa, (cudaStream_t)stream);
}

int tmg_squeeze_reverse(const float* y, float* x, int B, int C4, int H2, int W2, void* stream) {
  if (!x || !y) { set_error("null argument"); return TMG_ERR_NULL; }
  if (C4 < 4 || C4 % 4) { set_error("unsqueeze: %d channels not divisible by 4", C4); return TMG_ERR_BAD_SHAPE; }   // flowUtils.py:136
  PermArgs pa{};
  pa.src = y; pa.dst = x; pa.mode = PERM_UNSQUEEZE_NCHW_TO_NCHW; pa.B = B; pa.C = C4 / 4; pa.H = 2 * H2; pa.W = 2 * W2;
  return launch_permute(pa, (cudaStream_t)stream);
}

int tmg_nchw_to_nhwc(const float* src, float* dst, int B, int C, int H, int W, void* stream) {
  if (!src || !dst) { set_error("null argument"); return TMG_ERR_NULL; }
  PermArgs pa{};
  pa.src = src; pa.dst = dst; pa.mode = PERM_NCHW_TO_NHWC; pa.B = B; pa.C = C; pa.H = H; pa.W = W; pa.dst_cstride = C;
  return launch_permute(pa, (cudaStream_t)stream);
}

int tmg_nhwc_to_nchw(const float* src, float* dst, int B, int C, int H, int W, void* stream) {
  if (!src || !dst) { set_error("null argument"); return TMG_ERR_NULL; }
  PermArgs pa{};
  pa.src = src; pa.dst = dst; pa.mode = PERM_NHWC_TO_NCHW; pa.B = B; pa.C = C; pa.H = H; pa.W = W; pa.src_cstride = C;
  return launch_permute(pa, (cudaStream_t)stream);
}

size_t tmg_conv3x3_workspace_bytes(int Cin, int Cout) {
  size_t a = (size_t)9 * Cin * ((Cout + 3) / 4 * 4);
  size_t b = tc_packed_floats(Cin, tc_npad(Cout));
  return (a + b + 64) * sizeof(float) + sizeof(PackJob) + 256;
}

// Conv2d(k=3, s=1, p=1) (+ optional input ReLU, zero/replicate padding, bias, activation) on NHWC
// tensors with OIHW weights: the building block behind every conv of the path, exposed so the
// tensor-core kernel can be tested in isolation against the CUDA-core kernel and the oracle.
int tmg_conv3x3(int mode, const float* x, int B, int H, int W, int Cin, const float* w_oihw, const float* bias,
                int Cout, int relu_in, int pad_replicate, int act, float* out, void* workspace,
                size_t workspace_bytes, void* stream) {
  if (!x || !w_oihw || !out || !workspace) { set_error("null argument"); return TMG_ERR_NULL; }
  if (workspace_bytes < tmg_conv3x3_workspace_bytes(Cin, Cout)) { set_error("workspace too small"); return TMG_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = (float*)workspace;
  ConvSrc src{x, Cin, 0, Cin, relu_in};
  if (mode == TMG_PREC_FP32) {
    const int OP = (Cout + 3) / 4 * 4;
    float* wp = ws;
    PackJob j{};
    j.type = JOB_CONVW; j.a = Cout; j.b = Cin; j.opad = OP;
    for (auto& s : j.src) s = -1;
    j.src[0] = 0; j.dst[0] = 0;
    PackJob* jd = (PackJob*)(ws + (size_t)9 * Cin * OP + tc_packed_floats(Cin, tc_npad(Cout)) + 64);
    TMG_CUDA_OK(cudaMemcpyAsync(jd, &j, sizeof(j), cudaMemcpyHostToDevice, st));
    TMG_TRY(launch_pack(jd, 1, w_oihw, wp, 4, st));
    ConvArgs a{};
    a.src[0] = src; a.nsrc = 1; a.w = wp; a.cin_w = Cin; a.cout_w = OP; a.bias = bias; a.act = act;
    a.out = out; a.out_cstride = Cout; a.cout = Cout; a.B = B; a.Hin = H; a.Win = W; a.Hout = H; a.Wout = W;
    a.stride = 1; a.pad_replicate = pad_replicate;
    return launch_conv3x3(a, st);
  }
  if (mode != TMG_PREC_TF32X3 && mode != TMG_PREC_TF32) { set_error("unknown mode %d", mode); return TMG_ERR_BAD_CONFIG; }
  const int NP = tc_npad(Cout);
  if (NP > 256) { set_error("Cout %d > 256", Cout); return TMG_ERR_UNSUPPORTED; }
  float* wp = ws + (size_t)9 * Cin * ((Cout + 3) / 4 * 4);
  wp = (float*)(((uintptr_t)wp + 127) & ~(uintptr_t)127);
  TMG_TRY(launch_pack_tc(w_oihw, wp, Cout, Cin, NP, st));
  TcConvArgs t{};
  t.src[0] = src; t.nsrc = 1; t.cin = Cin; t.wpk = wp; t.npad = NP; t.bias = bias; t.act = act;
  t.out = out; t.out_cstride = Cout; t.cout = Cout; t.B = B; t.H = H; t.W = W; t.pad_replicate = pad_replicate;
  t.split3 = mode == TMG_PREC_TF32X3 ? 1 : 0;
  return launch_conv3x3_tc(t, st);
}

// Backward of nn.Conv2d(Cin, Cout, 3, padding=1) (+ input ReLU, zero / replicate padding) on NHWC tensors: data
// gradient, weight gradient (OIHW) and bias gradient.  Building block of the training path (autograd of every conv of
// the flow, nn/trainFlowParallel.py:277 loss.backward()); exposed for gradient parity tests.
size_t tmg_conv3x3_backward_workspace_bytes(int B, int H, int W, int Cin, int Cout) {
  return ((size_t)9 * Cout * ((Cin + 3) / 4 * 4) + wgrad_scratch_floats(Cout, Cin, B, H, W) + 128) * sizeof(float);
}

int tmg_conv3x3_backward(const float* x, int B, int H, int W, int Cin, const float* w_oihw, int Cout, int relu_in,
                         int pad_replicate, const float* gout, float* gx, float* gw, float* gbias, void* workspace,
                         size_t workspace_bytes, void* stream) {
  if (!x || !w_oihw || !gout || !workspace) { set_error("null argument"); return TMG_ERR_NULL; }
  if (workspace_bytes < tmg_conv3x3_backward_workspace_bytes(B, H, W, Cin, Cout)) { set_error("workspace too small"); return TMG_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = (float*)workspace;
  const int Ip = (Cin + 3) / 4 * 4;
  float* wt = ws;
  float* scratch = ws + (size_t)9 * Cout * Ip + 64;
  scratch = (float*)(((uintptr_t)scratch + 15) & ~(uintptr_t)15);
  if (gw) {
    WgradArgs wa{};
    wa.src[0] = ConvSrc{x, Cin, 0, Cin, relu_in}; wa.nsrc = 1; wa.cin = Cin;
    wa.g = gout; wa.g_cstride = Cout; wa.g_coff = 0; wa.cout = Cout;
    wa.B = B; wa.H = H; wa.W = W; wa.pad_replicate = pad_replicate;
    wa.gw = gw; wa.gbias = gbias; wa.accum = 0; wa.scratch = scratch;
    TMG_TRY(launch_wgrad(wa, st));
  }
  if (gx) {
    TMG_TRY(launch_pack_dgrad(w_oihw, wt, Cout, Cin, st));
    ConvArgs a{};
    a.src[0] = ConvSrc{gout, Cout, 0, Cout, 0}; a.nsrc = 1;
    a.w = wt; a.cin_w = Cout; a.cout_w = Ip; a.cout = Cin;
    a.out = gx; a.out_cstride = Cin; a.out_coff = 0;
    a.B = B; a.Hin = H; a.Win = W; a.Hout = H; a.Wout = W; a.stride = 1;
    a.mask = relu_in ? x : nullptr; a.accum = 0;
    TMG_TRY(launch_conv3x3(a, st));
    if (pad_replicate) {
      RingArgs r{};
      r.g = gout; r.g_cstride = Cout; r.g_coff = 0; r.cout = Cout;
      r.w_oihw = w_oihw; r.cin_total = Cin;
      r.ndst = 1; r.dst[0] = ConvDst{gx, relu_in ? x : nullptr, Cin, 0, Cin, 1};
      r.B = B; r.H = H; r.W = W;
      TMG_TRY(launch_dgrad_ring(r, st));
    }
  }
  return TMG_OK;
}

// Weight / bias gradient of the same convolution through the tensor-core kernel (wgrad_f16.cu, fp16 hi/lo split): what the
// training path runs in the f16x3 mode.  workspace: tmg_conv3x3_wgrad_tc_workspace_bytes.
size_t tmg_conv3x3_wgrad_tc_workspace_bytes(int B, int H, int W, int Cin, int Cout) {
  const int nch[1] = {Cin};
  return (wgrad_f16_scratch_floats(Cout, nch, 1, B, H, W) + 2048) * sizeof(float);
}
int tmg_conv3x3_wgrad_tc(const float* x, int B, int H, int W, int Cin, int Cout, int relu_in, int pad_replicate,
                         const float* gout, float* gw, float* gbias, void* workspace, size_t workspace_bytes, void* stream) {
  if (!x || !gout || !gw || !workspace) { set_error("null pointer"); return TMG_ERR_NULL; }
  if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) { set_error("bad shape"); return TMG_ERR_BAD_SHAPE; }
  cudaStream_t st = (cudaStream_t)stream;
  float* wsf = (float*)workspace;
  WgradArgs wa{};
  wa.src[0] = ConvSrc{x, Cin, 0, Cin, relu_in}; wa.nsrc = 1; wa.cin = Cin;
  wa.g = gout; wa.g_cstride = Cout; wa.g_coff = 0; wa.cout = Cout;
  wa.B = B; wa.H = H; wa.W = W; wa.pad_replicate = pad_replicate;
  wa.gw = gw; wa.gbias = gbias; wa.accum = 0; wa.scratch = wsf + 2048;
  if (!wgrad_f16_supported(wa)) { set_error("tensor-core weight gradient: unsupported shape Cin=%d Cout=%d", Cin, Cout); return TMG_ERR_UNSUPPORTED; }
  if (workspace_bytes < tmg_conv3x3_wgrad_tc_workspace_bytes(B, H, W, Cin, Cout)) { set_error("workspace too small"); return TMG_ERR_WORKSPACE; }
  TMG_TRY(launch_absmax_scale(gout, (int64_t)B * H * W, Cout, 0, Cout, wsf, wsf + 64, st));
  return launch_wgrad_f16(wa, wsf, st);
}

// Plan for the single-operator entry points: the workspace is sized by tmg_workspace_bytes of
// an LF input whose level-`level` flow map is Hl x Wl.
static int op_plan(tmg_model* m, int level, int B, int Hl, int Wl, Plan& p) {
  if (!m) { set_error("null model"); return TMG_ERR_NULL; }
  if (level < 0 || level >= m->cfg.n_levels) { set_error("bad level %d", level); return TMG_ERR_BAD_SHAPE; }
  int H = Hl << (level + 1), W = Wl << (level + 1);
  int up = m->cfg.cglow_upscale;
  if (H % up || W % up) { set_error("flow map %dx%d incompatible with upscale %d", Hl, Wl, up); return TMG_ERR_BAD_SHAPE; }
  return make_plan(*m, B, H / up, W / up, p);
}

int tmg_flow_step(tmg_model* m, int level, int step, int reverse, int B, int Hl, int Wl, const float* x,
                  const float* cond, const float* h_in, const float* c_in, float* out, float* logdet,
                  float* h_out, float* c_out, void* workspace, size_t workspace_bytes, void* stream) {
  Plan p;
  TMG_TRY(op_plan(m, level, B, Hl, Wl, p));
  TMG_TRY(check_common(m, workspace, workspace_bytes, p));
  if (!x || !cond || !out || !logdet) { set_error("null argument"); return TMG_ERR_NULL; }
  const LevelW& lv = m->levels[level];
  if (step < 1 || step > (int)lv.steps.size()) { set_error("bad step %d", step); return TMG_ERR_BAD_SHAPE; }
  const StepW& st = lv.steps[step - 1];
  Ctx c{*m, p, (float*)workspace, (cudaStream_t)stream};
  float* ws = c.ws;
  const int HW = Hl * Wl, C = lv.C, cf = m->cfg.cond_features;
  float* Y = ws + p.scratch_in;
  float* CN = ws + p.scratch_cond;
  const int ldstride = p.nslots * p.ctas;
  TMG_CUDA_OK(cudaMemsetAsync(ws + p.ldp, 0, (size_t)B * ldstride * sizeof(float), c.st));
  PermArgs pa{};
  pa.src = x; pa.dst = Y; pa.mode = PERM_NCHW_TO_NHWC; pa.B = B; pa.C = C; pa.H = Hl; pa.W = Wl; pa.dst_cstride = C;
  TMG_TRY(launch_permute(pa, c.st));
  pa.src = cond; pa.dst = CN; pa.C = cf; pa.dst_cstride = cf;
  TMG_TRY(launch_permute(pa, c.st));
  float* Y2 = ws + p.scratch_out;
  if (!reverse) TMG_TRY(run_pointwise(c, level, Y, false, &st, false, B, HW, nullptr));
  TMG_TRY(run_step(c, level, st, reverse ? &st : nullptr, reverse != 0, B, Hl, Wl, Y, Y2, CN, h_in, c_in, h_out, c_out,
                   ws + p.ldp));
  PermArgs pb{};
  pb.src = Y; pb.dst = out; pb.mode = PERM_NHWC_TO_NCHW; pb.B = B; pb.C = C; pb.H = Hl; pb.W = Wl; pb.src_cstride = C;
  TMG_TRY(launch_permute(pb, c.st));
  // log-det of this single step: partial sums + (sum log|w| - sum log_s) * HW
  LogdetArgs a{};
  a.ld_part = ws + p.ldp; a.ld_stride = ldstride;
  a.step_const = c.Q() + m->step_const_off;
  a.n_levels = 1; a.step_begin[0] = st.const_idx; a.step_begin[1] = st.const_idx + 1; a.hw[0] = HW;
  a.out = logdet; a.B = B;
  return launch_logdet_reduce(a, c.st);
}

// ---- backward of one REVERSE flow step: gradients w.r.t. the step input, the conditioning map, the incoming LSTM
// states and every parameter of the step (accumulated into a flat gradient buffer laid out like the parameter
// buffer).  The forward is recomputed with the exact-fp32 kernels.
struct BwdExtra {
  size_t go, gy, gz, gu, v, gcond, gd, part, dw, tmp, wt, wscr, oscr, gscale;
  size_t hn, cn, ghn, ggates, gu0;      // LSTM step
  size_t total;
};

static BwdExtra bwd_extra(const tmg_model& m, int level, int B, int Hl, int Wl) {
  const int C = m.levels[level].C, cf = m.cfg.cond_features, cin_t = C / 2 + cf, R = m.cfg.rec_features;
  const size_t px = (size_t)B * Hl * Wl;
  BwdExtra e{};
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += align_up(n, 64); return o; };
  e.go = take(px * C); e.gy = take(px * C); e.gz = take(px * C); e.gu = take(px * C); e.v = take(px * C);
  e.gcond = take(px * cf); e.gd = take(px * 2);
  e.part = take((size_t)step_bwd_blocks(B, Hl * Wl) * (3 * C + 2));
  e.dw = take((size_t)C * C); e.tmp = take(16);
  // transposed weights of the widest convolution (gate conv) and the largest weight-gradient scratch
  size_t wt = (size_t)9 * std::max(C, 4 * R) * ((cin_t + R + 3) / 4 * 4) + 64;
  e.wt = take(wt);
  size_t ws = wgrad_scratch_floats(C, cin_t + 2, B, Hl, Wl);
  ws = std::max(ws, wgrad_scratch_floats(1, cin_t + 1, B, Hl, Wl));
  ws = std::max(ws, wgrad_scratch_floats(4 * R, cin_t + R, B, Hl, Wl));
  ws = std::max(ws, wgrad_scratch_floats(cin_t, cin_t + R, B, Hl, Wl));
  {   // tensor-core weight gradients (f16 modes)
    const int n_zc[3] = {C / 2, cf, 2}, n_zl[2] = {cin_t, 2}, n_g[3] = {C / 2, cf, R}, n_d2[3] = {C / 2, cf, 1}, n_d2l[2] = {cin_t, 1};
    ws = std::max(ws, wgrad_f16_scratch_floats(C, n_zc, 3, B, Hl, Wl));
    ws = std::max(ws, wgrad_f16_scratch_floats(C, n_zl, 2, B, Hl, Wl));
    ws = std::max(ws, wgrad_f16_scratch_floats(4 * R, n_g, 3, B, Hl, Wl));
    ws = std::max(ws, wgrad_f16_scratch_floats(cin_t, n_g, 3, B, Hl, Wl));
    ws = std::max(ws, wgrad_f16_scratch_floats(1, n_d2, 3, B, Hl, Wl));
    ws = std::max(ws, wgrad_f16_scratch_floats(1, n_d2l, 2, B, Hl, Wl));
    ws = std::max(ws, wgrad_cout1_scratch_floats(cin_t + 1, B, Hl, Wl));
    const int n_sp[1] = {C / 2};
    ws = std::max(ws, wgrad_f16_scratch_floats(C, n_sp, 1, B, Hl, Wl));
  }
  e.wscr = take(ws);
  e.oscr = take(outer_wgrad_scratch_floats((int64_t)px, C));
  e.gscale = take(1024 + 64);           // [0..1] power-of-two scale of a gradient tensor and its inverse, [64..] block maxima
  e.hn = take(px * R); e.cn = take(px * R); e.ghn = take(px * R); e.ggates = take(px * 4 * R);
  e.gu0 = take(px * ((cin_t + 3) / 4 * 4));
  e.total = off;
  return e;
}

// weight gradient (accumulated into the flat gradient buffer) + data gradient of one 3x3 convolution into up to three
// destination slices (the sources of the virtual concatenation), with ReLU gating and replicate-padding border terms
struct BwdDest { float* g; const float* fwd; int cstride, coff, nch; int accum; };
static int conv_backward(Ctx& c, int B, int Hl, int Wl, const ConvW& w, int nsrc_fwd, const ConvSrc* fsrc, bool replicate,
                         const float* g, int g_cs, int g_co, const BwdDest* dests, int ndest, float* grads, float* wt,
                         float* wscr, float* gscale = nullptr, bool pre_scaled = false) {
  WgradArgs wa{};
  for (int i = 0; i < nsrc_fwd; ++i) wa.src[i] = fsrc[i];
  wa.nsrc = nsrc_fwd; wa.cin = w.I;
  wa.g = g; wa.g_cstride = g_cs; wa.g_coff = g_co; wa.cout = w.O;
  wa.B = B; wa.H = Hl; wa.W = Wl; wa.pad_replicate = replicate ? 1 : 0;
  wa.gw = grads + w.w_param; wa.gbias = w.b_param >= 0 ? grads + w.b_param : nullptr; wa.accum = 1; wa.scratch = wscr;
  // f16 modes: weight gradient (pixels as the GEMM K dimension, MN-major operands) and data gradient on tensor cores;
  // the output gradient is scaled into the fp16 range by a power of two measured once per convolution
  const bool tc_bwd = gscale && prec_f16(c.m.precision);
  static const bool wg_off = [] { const char* e = getenv("TMG_WGRAD_FFMA"); return e && e[0] == '1'; }();
  if (gscale && ndest <= 3 && wgrad_cout1_supported(wa)) {
    // dense layers (Cout = 1): both adjoints are streaming kernels in exact fp32 -- on the tensor cores they are an
    // M = 1 / K = 1 GEMM padded to a full tile, i.e. pure pipeline latency; no gradient scaling needed either
    ConvDst cd[3];
    int tot = 0;
    for (int d = 0; d < ndest; ++d) {
      cd[d] = ConvDst{dests[d].g, dests[d].fwd, dests[d].cstride, dests[d].coff, dests[d].nch, dests[d].accum};
      tot += dests[d].nch;
    }
    if (tot == w.I) {
      TMG_TRY(launch_wgrad_cout1(wa, c.st));
      return launch_dgrad_cout1(g, g_cs, g_co, c.P() + w.w_param, w.I, cd, ndest, B, Hl, Wl, c.st);
    }
  }
  const bool wg_tc = tc_bwd && !wg_off && wgrad_f16_supported(wa);
  if (pre_scaled) {
    wa.gbias = nullptr;          // the caller's reduction already produced the bias gradient and filled gscale
  } else if (wg_tc && wa.gbias && w.O <= 256) {
    // the scale of g and the bias gradient (column sums of g) in one pass over g; the column-sum partials live at the
    // tail of the weight-gradient scratch (wgrad_f16_scratch_floats reserves them)
    float* pb = wgrad_f16_bias_partials(wa);
    TMG_TRY(launch_absmax_colsum(g, (int64_t)B * Hl * Wl, g_cs, g_co, w.O, gscale, c.m.sync_dev, pb, wa.gbias, 1, c.st));
    wa.gbias = nullptr;
  } else if (tc_bwd) {
    TMG_TRY(launch_absmax_scale(g, (int64_t)B * Hl * Wl, g_cs, g_co, w.O, gscale, gscale + 64, c.st, c.m.sync_dev));
  }
  if (wg_tc) TMG_TRY(launch_wgrad_f16(wa, gscale, c.st));
  else TMG_TRY(launch_wgrad(wa, c.st));
  // data gradient on tensor cores (f16 modes): ONE launch of the fp16 hi/lo conv kernel on the transposed, tap-flipped
  // weights with the output gradient scaled into the fp16 range; its epilogue routes the columns to the destinations
  bool dgrad_tc = false;
  if (gscale && w.O == 1 && !replicate && ndest <= 3) {
    // dense layers (Cout = 1): nine FMAs per element, exact fp32 streaming kernel instead of a K = 1 GEMM
    ConvDst cd[3];
    int tot = 0;
    for (int d = 0; d < ndest; ++d) {
      cd[d] = ConvDst{dests[d].g, dests[d].fwd, dests[d].cstride, dests[d].coff, dests[d].nch, dests[d].accum};
      tot += dests[d].nch;
    }
    if (tot == w.I) {
      TMG_TRY(launch_dgrad_cout1(g, g_cs, g_co, c.P() + w.w_param, w.I, cd, ndest, B, Hl, Wl, c.st));
      return TMG_OK;
    }
  }
  if (gscale && prec_f16(c.m.precision) && w.w_pack_f16t >= 0 && ndest <= 3) {
    if (!tc_bwd && !pre_scaled) TMG_TRY(launch_absmax_scale(g, (int64_t)B * Hl * Wl, g_cs, g_co, w.O, gscale, gscale + 64, c.st, c.m.sync_dev));
    ConvF16Args t{};
    t.src[0] = ConvSrc{g, g_cs, g_co, w.O, 0}; t.nsrc = 1;
    t.wpk = c.Q() + w.w_pack_f16t; t.inv_scale = c.Q() + w.inv_f16t; t.npad = w.NPt; t.cout = w.I;
    t.B = B; t.H = Hl; t.W = Wl; t.pad_replicate = 0; t.x3 = prec_split(c.m.precision) ? 1 : 0; t.overflow = c.m.sync_dev ? c.m.sync_dev + 32 : nullptr;
    t.in_scale = gscale;
    t.ndst = ndest;
    int tot = 0;
    bool split_ok = true;          // the destinations are the forward sources the transposed weights were packed for
    for (int d = 0; d < 3; ++d) {
      if (d < ndest) {
        t.dst[d] = ConvDst{dests[d].g, dests[d].fwd, dests[d].cstride, dests[d].coff, dests[d].nch, dests[d].accum};
        tot += dests[d].nch;
      }
      split_ok = split_ok && (d < ndest ? dests[d].nch : 0) == w.f16_nch[d];
    }
    if (tot == w.I && split_ok && convf16_supported(t)) {
      TMG_TRY(launch_conv3x3_f16(t, c.st));
      dgrad_tc = true;
    }
  }
  if (!dgrad_tc) TMG_TRY(launch_pack_dgrad(c.P() + w.w_param, wt, w.O, w.I, c.st));
  const int Ip = (w.I + 3) / 4 * 4;
  int c0 = 0;
  for (int d = 0; d < ndest; ++d) {
    if (dests[d].g) {
      ConvArgs a{};
      a.src[0] = ConvSrc{g, g_cs, g_co, w.O, 0}; a.nsrc = 1;
      a.w = wt + c0; a.cin_w = w.O; a.cout_w = Ip; a.cout = dests[d].nch;
      a.out = dests[d].g; a.out_cstride = dests[d].cstride; a.out_coff = dests[d].coff;
      a.B = B; a.Hin = Hl; a.Win = Wl; a.Hout = Hl; a.Wout = Wl; a.stride = 1;
      a.mask = dests[d].fwd; a.accum = dests[d].accum;
      if (!dgrad_tc) TMG_TRY(launch_conv3x3(a, c.st));
    }
    c0 += dests[d].nch;
  }
  if (replicate && ndest <= 3 && c0 == w.I) {     // border terms of replicate padding, all destinations in one launch
    RingArgs r{};
    r.g = g; r.g_cstride = g_cs; r.g_coff = g_co; r.cout = w.O;
    r.w_oihw = c.P() + w.w_param; r.cin_total = w.I;
    r.ndst = ndest;
    for (int d = 0; d < ndest; ++d)
      r.dst[d] = ConvDst{dests[d].g, dests[d].fwd, dests[d].cstride, dests[d].coff, dests[d].nch, 1};
    r.B = B; r.H = Hl; r.W = Wl;
    TMG_TRY(launch_dgrad_ring(r, c.st));
  } else if (replicate) {
    set_error("conv_backward: replicate-padding destinations do not cover the input channels");
    return TMG_ERR_UNSUPPORTED;
  }
  return TMG_OK;
}

struct StepBwdIO {
  const float* Y;        // step input, NHWC [B,HW,C]
  const float* COND;     // NHWC [B,HW,cond]
  const float* GO;       // gradient w.r.t. the step output, NHWC
  const float* g_ld;     // [B]
  float* GY;             // out: gradient w.r.t. Y (written)
  float* GC;             // gradient w.r.t. COND: ACCUMULATED (the caller zeroes it)
  const float* h_prev; const float* c_prev;     // LSTM step: incoming states (null = zeros)
  const float* g_hn; const float* g_cn;         // gradients w.r.t. the returned states (null = zero)
  float* g_hprev; float* g_cprev;               // out (written when non-null)
  float* grads;
  bool defer_lu = false;            // accumulate dW and hw*sum(g_ld) into the stash slots; tmg_backward_finalize finishes
  const float* D_tape = nullptr;    // relu(d1), relu(d2) and h recorded by the training forward (plain steps), else null:
  const float* HR_tape = nullptr;   // the coupling network is then recomputed from Y
};

static int step_backward(Ctx& c, int level, const StepW& st, int B, int Hl, int Wl, const StepBwdIO& io, float* ex,
                         const BwdExtra& e) {
  tmg_model* m = &c.m;
  const Plan& p = c.p;
  float* ws = c.ws;
  const LevelW& lv = m->levels[level];
  const int HW = Hl * Wl, C = lv.C, cf = m->cfg.cond_features, cin_t = C / 2 + cf, R = m->cfg.rec_features;
  const int u0s = (cin_t + 3) / 4 * 4;
  const float* Y = io.Y; const float* CN = io.COND;
  const bool lstm = st.kind == STEP_LSTM;
  const bool taped = !lstm && io.D_tape && io.HR_tape;
  const float* D = taped ? io.D_tape : ws + p.d;
  const float* HR = taped ? io.HR_tape : ws + p.hr;
  float *GY = io.GY, *GC = io.GC, *GZ = ex + e.gz, *GU = ex + e.gu, *V = ex + e.v, *GD = ex + e.gd;
  float *HN = ex + e.hn, *CNW = ex + e.cn, *GHN = ex + e.ghn, *GG = ex + e.ggates, *GU0 = ex + e.gu0, *U0 = ws + p.u0;
  float* grads = io.grads;
  // forward recompute with the exact-fp32 kernels: D (d1, d2), HR (h) and, for the LSTM step, gates / h' / c' / u0
  // (f16x3 / f16 modes: every conv on its own through conv3x3_f16_kernel, fp32-grade in f16x3; else the fp32 kernels)
  const int prec = m->precision;
  const bool was_unfused = c.unfused;
  c.unfused = true;
  if (!prec_f16(prec)) m->precision = TMG_PREC_FP32;
  int rc = taped ? TMG_OK : run_coupling_nn(c, level, st, B, Hl, Wl, Y, CN, io.h_prev, io.c_prev, lstm ? HN : nullptr, lstm ? CNW : nullptr);
  m->precision = prec;
  c.unfused = was_unfused;
  TMG_TRY(rc);
  const bool normed = st.kind != STEP_UNNORMED;
  StepBwdArgs sa{};
  sa.y_in = Y; sa.h = HR; sa.g_out = io.GO; sa.g_ld = io.g_ld;
  sa.wmat = c.Q() + st.W;
  if (normed) { sa.nw = c.P() + st.norm_w; sa.nb = c.P() + st.norm_b; }
  sa.gain = c.Q() + st.zc_gain;
  sa.g_y = GY; sa.g_z = GZ; sa.gu = GU; sa.v = V; sa.part = ex + e.part;
  sa.B = B; sa.HW = HW; sa.C = C;
  TMG_TRY(launch_step_bwd(sa, c.st));
  const int nblk = step_bwd_blocks(B, HW), pstride = 3 * C + 2;
  // f16 modes: the bias gradient of the Conv2dZeros conv and the scale of its output gradient come out of this reduction
  const bool pre = prec_f16(m->precision) && st.zc.b_param >= 0;
  (void)pstride;
  TMG_TRY(launch_step_param_grads(ex + e.part, nblk, C, normed ? grads + st.norm_b : nullptr, normed ? grads + st.norm_w : nullptr,
                                  c.P() + st.zc_scale, grads + st.zc_scale, io.g_ld, B, (float)HW,
                                  io.defer_lu ? grads + st.lu[4] : nullptr, pre ? grads + st.zc.b_param : nullptr,
                                  pre ? ex + e.gscale : nullptr, c.st));
  // 1x1 convolution: dW, then the LU parameterisation and the log-det constants
  if (io.defer_lu) {
    // linear in dW and in hw * sum(g_ld): accumulated over the time steps of a BPTT block in the gradient slots of the
    // buffers `p` / `sign_s` and turned into the gradients of l, u, log_s, norm.weight once (tmg_backward_finalize)
    TMG_TRY(launch_outer_wgrad(GU, V, (int64_t)B * HW, C, grads + st.lu[3], ex + e.oscr, c.st, 1));
  } else {
  TMG_TRY(launch_outer_wgrad(GU, V, (int64_t)B * HW, C, ex + e.dw, ex + e.oscr, c.st));
  LuBwdArgs la{};
  la.dW = ex + e.dw;
  la.l = c.P() + st.lu[0]; la.u = c.P() + st.lu[1]; la.log_s = c.P() + st.lu[2]; la.p = c.P() + st.lu[3];
  la.sign_s = c.P() + st.lu[4]; la.lmask = c.P() + st.lu[5]; la.umask = c.P() + st.lu[6]; la.eye = c.P() + st.lu[7];
  la.g_l = grads + st.lu[0]; la.g_u = grads + st.lu[1]; la.g_log_s = grads + st.lu[2];
  if (normed) { la.nw = c.P() + st.norm_w; la.g_nw = grads + st.norm_w; }
  la.g_ld = io.g_ld; la.B = B; la.hw = (float)HW; la.C = C;
  TMG_TRY(launch_lu_bwd(la, c.st));
  }

  auto conv_bwd = [&](const ConvW& w, int nsrc_fwd, const ConvSrc* fsrc, bool replicate, const float* g, int g_cs, int g_co,
                      const BwdDest* dests, int ndest, bool pre_scaled = false) -> int {
    return conv_backward(c, B, Hl, Wl, w, nsrc_fwd, fsrc, replicate, g, g_cs, g_co, dests, ndest, grads, ex + e.wt, ex + e.wscr,
                         ex + e.gscale, pre_scaled);
  };
  typedef BwdDest Dest;

  if (!lstm) {
    // coupling network on t = cat(y1, cond): three convolutions, last to first
    const ConvSrc f3[3] = {{Y, C, 0, C / 2, 1}, {CN, cf, 0, cf, 1}, {D, 2, 0, 2, 1}};
    const Dest d3[3] = {{GY, Y, C, 0, C / 2, 1}, {GC, CN, cf, 0, cf, 1}, {GD, D, 2, 0, 2, 0}};
    TMG_TRY(conv_bwd(st.zc, 3, f3, true, GZ, C, 0, d3, 3, pre));
    const ConvSrc f2[3] = {{Y, C, 0, C / 2, 1}, {CN, cf, 0, cf, 1}, {D, 2, 0, 1, 1}};
    const Dest d2[3] = {{GY, Y, C, 0, C / 2, 1}, {GC, CN, cf, 0, cf, 1}, {GD, D, 2, 0, 1, 1}};
    TMG_TRY(conv_bwd(st.d2, 3, f2, false, GD, 2, 1, d2, 3));
    const Dest d1[2] = {{GY, Y, C, 0, C / 2, 1}, {GC, CN, cf, 0, cf, 1}};
    TMG_TRY(conv_bwd(st.d1, 2, f3, false, GD, 2, 0, d1, 2));
    return TMG_OK;
  }
  // ---- LSTM step (flowAffine.py:161-236, convLSTM.py:54-152): coupling network on u0 = relu(LSTM_out_conv(cat(t, h')))
  {
    const ConvSrc f3[2] = {{U0, u0s, 0, cin_t, 1}, {D, 2, 0, 2, 1}};
    const Dest d3[2] = {{GU0, U0, u0s, 0, cin_t, 0}, {GD, D, 2, 0, 2, 0}};
    TMG_TRY(conv_bwd(st.zc, 2, f3, true, GZ, C, 0, d3, 2, pre));
    const ConvSrc f2[2] = {{U0, u0s, 0, cin_t, 1}, {D, 2, 0, 1, 1}};
    const Dest d2[2] = {{GU0, U0, u0s, 0, cin_t, 1}, {GD, D, 2, 0, 1, 1}};
    TMG_TRY(conv_bwd(st.d2, 2, f2, false, GD, 2, 1, d2, 2));
    const Dest d1[1] = {{GU0, U0, u0s, 0, cin_t, 1}};
    TMG_TRY(conv_bwd(st.d1, 1, f3, false, GD, 2, 0, d1, 1));
  }
  // GU0 is gated by u0 > 0, i.e. it is already the gradient w.r.t. the pre-ReLU output of LSTM_out_conv
  if (io.g_hn) TMG_CUDA_OK(cudaMemcpyAsync(GHN, io.g_hn, (size_t)B * HW * R * sizeof(float), cudaMemcpyDeviceToDevice, c.st));
  else TMG_CUDA_OK(cudaMemsetAsync(GHN, 0, (size_t)B * HW * R * sizeof(float), c.st));
  {
    const ConvSrc fo[3] = {{Y, C, 0, C / 2, 0}, {CN, cf, 0, cf, 0}, {HN, R, 0, R, 0}};
    const Dest dox[3] = {{GY, nullptr, C, 0, C / 2, 1}, {GC, nullptr, cf, 0, cf, 1}, {GHN, nullptr, R, 0, R, 1}};
    TMG_TRY(conv_bwd(st.outc, 3, fo, false, GU0, u0s, 0, dox, 3));
  }
  LstmBwdArgs lb{};
  lb.gates = ws + p.gates; lb.c_prev = io.c_prev; lb.g_h = GHN; lb.g_c = io.g_cn;
  lb.g_gates = GG; lb.g_c_prev = io.g_cprev; lb.n = (int64_t)B * HW * R; lb.R = R;
  TMG_TRY(launch_lstm_bwd(lb, c.st));
  {
    const ConvSrc fg[3] = {{Y, C, 0, C / 2, 0}, {CN, cf, 0, cf, 0}, {io.h_prev, R, 0, R, 0}};
    const Dest dg[3] = {{GY, nullptr, C, 0, C / 2, 1}, {GC, nullptr, cf, 0, cf, 1}, {io.g_hprev, nullptr, R, 0, R, 0}};
    TMG_TRY(conv_bwd(st.gate, 3, fg, false, GG, 4 * R, 0, dg, 3));
  }
  return TMG_OK;
}

// ---- backward of TMGlow.reconstruct / TMGlow.sample (tmGlow.py:417-467 + LSTMCFlowDecoder.reverse :269-303) given the
// tape of the training forward.  Decoder (flow) parameters, LSTM state gradients and the gradient w.r.t. the
// conditioning maps / top prior parameters; the encoder backward follows (run_encoder_backward).
struct RbExtra { BwdExtra e; size_t ga, gb, gcond[TMG_MAX_LEVELS], gzout, gpart;
                 size_t gdb[TMG_MAX_LEVELS], gcc, gact, ge0, bnsums, encw, encwt; size_t total; };

static RbExtra rb_extra(const tmg_model& m, const Plan& p) {
  RbExtra r{};
  r.e = bwd_extra(m, 0, p.B, p.Hl[0], p.Wl[0]);
  for (int l = 1; l < p.L; ++l) {          // level 0 is the largest in pixels; wider levels need wider per-pixel rows
    BwdExtra q = bwd_extra(m, l, p.B, p.Hl[l], p.Wl[l]);
    r.e.total = std::max(r.e.total, q.total);
  }
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += align_up(n, 64); return o; };
  size_t mx = 0, mxpix = 0;
  for (int l = 0; l < p.L; ++l) { mx = std::max(mx, (size_t)p.B * p.Hl[l] * p.Wl[l] * m.levels[l].C); mxpix = std::max(mxpix, (size_t)p.B * p.Hl[l] * p.Wl[l]); }
  r.ga = take(mx); r.gb = take(mx);
  for (int l = 0; l < p.L; ++l) r.gcond[l] = take((size_t)p.B * p.Hl[l] * p.Wl[l] * m.cfg.cond_features);
  r.gzout = take((size_t)p.B * p.Hl[p.L - 1] * p.Wl[p.L - 1] * 2 * m.Cz);
  r.gpart = take((size_t)gauss_bwd_blocks(p.B, p.Hl[0] * p.Wl[0]) + 64);
  {   // encoder backward
    const tmg_config& g = m.cfg;
    size_t mxcc = 0, mxact = 0, encw = 0, encwt = 0;
    int nfmax = 0;
    auto need = [&](int O, int I, int H, int W) {
      encw = std::max(encw, wgrad_scratch_floats(O, I, p.B, H, W));
      encwt = std::max(encwt, (size_t)9 * O * ((I + 3) / 4 * 4) + 64);
    };
    for (int l = 0; l < p.L; ++l) {
      const LevelW& lv = m.levels[l];
      const size_t px = (size_t)p.B * p.eh[l] * p.ew[l];
      r.gdb[l] = take(px * lv.nf_out);
      mxcc = std::max(mxcc, px * std::max(g.cond_features, 2 * m.Cz));
      mxact = std::max(mxact, px * lv.nf_out);
      nfmax = std::max(nfmax, lv.nf_out);
      need(g.cond_features, lv.nf_out, p.eh[l], p.ew[l]);
      need(2 * m.Cz, lv.nf_out, p.eh[l], p.ew[l]);
      need(g.growth_rate, lv.nf_out, p.eh[l], p.ew[l]);
      need(lv.nf_in, l > 0 ? m.levels[l - 1].nf_out : g.init_features / 2, p.eh[l], p.ew[l]);
    }
    need(g.init_features / 2, g.in_features, p.h, p.w);
    r.gcc = take(mxcc); r.gact = take(mxact);
    r.ge0 = take((size_t)p.B * p.h * p.w * (g.init_features / 2));
    r.bnsums = take((size_t)2 * nfmax + 64);
    r.encw = take(encw); r.encwt = take(encwt);
  }
  r.total = off;
  return r;
}

// Encoder.forward backward (nn/tmGlow.py:104-186, denseBlock.py:15-100, misc.py:34): from the gradients w.r.t. the
// conditioning maps and the top prior parameters to every encoder parameter.  The activations are in the workspace
// (run_encoder was just re-run); BatchNorm uses batch statistics (train) or the running ones (eval).
static int run_encoder_backward(Ctx& c, bool bn_train, float* rb, const RbExtra& rx, float* grads) {
  tmg_model& m = c.m;
  const tmg_config& g = m.cfg;
  const Plan& p = c.p;
  float* ws = c.ws;
  const int B = p.B, L = p.L, up = g.cglow_upscale;
  float* wt = rb + rx.encwt;
  float* wscr = rb + rx.encw;
  float* GCC = rb + rx.gcc;
  float* GA = rb + rx.gact;
  for (int l = 0; l < L; ++l)
    TMG_CUDA_OK(cudaMemsetAsync(rb + rx.gdb[l], 0, (size_t)B * p.eh[l] * p.ew[l] * m.levels[l].nf_out * sizeof(float), c.st));
  for (int i = L - 1; i >= 0; --i) {
    const LevelW& lv = m.levels[i];
    const int eh = p.eh[i], ew = p.ew[i];
    const int64_t N = (int64_t)B * eh * ew;
    float* db = ws + p.db[i];
    float* GDB = rb + rx.gdb[i];
    const ConvSrc sdb{db, lv.nf_out, 0, lv.nf_out, 0};
    const BwdDest ddb{GDB, nullptr, lv.nf_out, 0, lv.nf_out, 1};
    if (i == L - 1) {
      const float* gz = rb + rx.gzout;
      if (up > 1) { TMG_TRY(launch_upsample_bwd(gz, GCC, B, eh, ew, 2 * m.Cz, up, c.st)); gz = GCC; }
      TMG_TRY(conv_backward(c, B, eh, ew, m.out_conv, 1, &sdb, false, gz, 2 * m.Cz, 0, &ddb, 1, grads, wt, wscr));
    }
    {
      const float* gc = rb + rx.gcond[i];
      if (up > 1) { TMG_TRY(launch_upsample_bwd(gc, GCC, B, eh, ew, g.cond_features, up, c.st)); gc = GCC; }
      TMG_TRY(conv_backward(c, B, eh, ew, lv.cond, 1, &sdb, false, gc, g.cond_features, 0, &ddb, 1, grads, wt, wscr));
    }
    for (int l = (int)lv.dense.size() - 1; l >= 0; --l) {
      const DenseW& d = lv.dense[l];
      const float *mean, *var, *sc, *sh;
      if (bn_train) {
        BnStatArgs sa{};
        sa.x = db; sa.cstride = lv.nf_out; sa.c0 = 0; sa.n = d.cin; sa.N = N; sa.mean = ws + p.bn_mean; sa.var = ws + p.bn_var;
        TMG_TRY(launch_bn_stats(sa, c.st));
        BnFoldArgs fa{};
        fa.mean = ws + p.bn_mean; fa.var = ws + p.bn_var; fa.w = c.P() + d.bn_w; fa.b = c.P() + d.bn_b;
        fa.run_mean = m.params + d.bn_rm; fa.run_var = m.params + d.bn_rv;
        fa.scale = ws + p.bn_scale; fa.shift = ws + p.bn_shift; fa.n = d.cin; fa.N = N; fa.eps = 1e-5f; fa.momentum = 0.f;
        TMG_TRY(launch_bn_fold_train(fa, c.st));
        mean = ws + p.bn_mean; var = ws + p.bn_var; sc = ws + p.bn_scale; sh = ws + p.bn_shift;
      } else {
        mean = c.P() + d.bn_rm; var = c.P() + d.bn_rv; sc = c.Q() + d.scale; sh = c.Q() + d.shift;
      }
      // conv: weight gradient on relu(bn(x)), data gradient w.r.t. relu(bn(x)) into GA
      WgradArgs wa{};
      wa.src[0] = ConvSrc{db, lv.nf_out, 0, d.cin, 1}; wa.nsrc = 1; wa.cin = d.cin;
      wa.bn_scale = sc; wa.bn_shift = sh;
      wa.g = GDB; wa.g_cstride = lv.nf_out; wa.g_coff = d.cin; wa.cout = d.conv.O;
      wa.B = B; wa.H = eh; wa.W = ew; wa.gw = grads + d.conv.w_param; wa.accum = 1; wa.scratch = wscr;
      TMG_TRY(launch_wgrad(wa, c.st));
      TMG_TRY(launch_pack_dgrad(c.P() + d.conv.w_param, wt, d.conv.O, d.conv.I, c.st));
      ConvArgs a{};
      a.src[0] = ConvSrc{GDB, lv.nf_out, d.cin, d.conv.O, 0}; a.nsrc = 1;
      a.w = wt; a.cin_w = d.conv.O; a.cout_w = (d.cin + 3) / 4 * 4; a.cout = d.cin;
      a.out = GA; a.out_cstride = d.cin; a.out_coff = 0;
      a.B = B; a.Hin = eh; a.Win = ew; a.Hout = eh; a.Wout = ew; a.stride = 1;
      TMG_TRY(launch_conv3x3(a, c.st));
      BnBwdArgs ba{};
      ba.x = db; ba.x_cstride = lv.nf_out; ba.ga = GA; ba.ga_cstride = d.cin; ba.mean = mean; ba.var = var;
      ba.gamma = c.P() + d.bn_w; ba.beta = c.P() + d.bn_b; ba.gx = GDB; ba.gx_cstride = lv.nf_out;
      ba.g_gamma = grads + d.bn_w; ba.g_beta = grads + d.bn_b; ba.sums = rb + rx.bnsums;
      ba.n = d.cin; ba.N = N; ba.eps = 1e-5f; ba.train = bn_train ? 1 : 0;
      TMG_TRY(launch_bn_relu_bwd(ba, c.st));
    }
    // the stride-2 convolution that produced the first nf_in channels of this level
    const ConvW& cw = i > 0 ? lv.trans : m.in_conv3;
    const float* src = i > 0 ? ws + p.db[i - 1] : ws + p.e0;
    const int sch = i > 0 ? m.levels[i - 1].nf_out : g.init_features / 2;
    const int Hin = i > 0 ? p.eh[i - 1] : p.h, Win = i > 0 ? p.ew[i - 1] : p.w;
    float* gsrc = i > 0 ? rb + rx.gdb[i - 1] : rb + rx.ge0;
    WgradArgs wa{};
    wa.src[0] = ConvSrc{src, sch, 0, sch, 1}; wa.nsrc = 1; wa.cin = sch;
    wa.g = GDB; wa.g_cstride = lv.nf_out; wa.g_coff = 0; wa.cout = cw.O;
    wa.B = B; wa.H = eh; wa.W = ew; wa.stride = 2; wa.Hin = Hin; wa.Win = Win;
    wa.gw = grads + cw.w_param; wa.accum = 1; wa.scratch = wscr;
    TMG_TRY(launch_wgrad(wa, c.st));
    S2DgradArgs da{};
    da.g = GDB; da.g_cstride = lv.nf_out; da.g_coff = 0; da.cout = cw.O; da.Hout = eh; da.Wout = ew;
    da.w_oihw = c.P() + cw.w_param; da.cin = sch; da.mask = src; da.gx = gsrc; da.gx_cstride = sch; da.gx_coff = 0;
    da.accum = i > 0 ? 1 : 0; da.B = B; da.Hin = Hin; da.Win = Win;
    TMG_TRY(launch_dgrad_s2(da, c.st));
  }
  // In_conv (tmGlow.py:148): weight gradient only (the LF input needs no gradient)
  WgradArgs wa{};
  wa.src[0] = ConvSrc{ws + p.xn, g.in_features, 0, g.in_features, 0}; wa.nsrc = 1; wa.cin = g.in_features;
  wa.g = rb + rx.ge0; wa.g_cstride = g.init_features / 2; wa.g_coff = 0; wa.cout = m.in_conv.O;
  wa.B = B; wa.H = p.h; wa.W = p.w; wa.gw = grads + m.in_conv.w_param; wa.accum = 1; wa.scratch = wscr;
  TMG_TRY(launch_wgrad(wa, c.st));
  return TMG_OK;
}

size_t tmg_reconstruct_backward_workspace_bytes(const tmg_model* m, int B, int h, int w) {
  if (!m) return 0;
  Plan p;
  if (make_plan(*m, B, h, w, p) != TMG_OK) return 0;
  const RbExtra r = rb_extra(*m, p);
  return align_up(p.total, 256) + (r.e.total + r.total) * sizeof(float) + 512;
}

static int reconstruct_backward_impl(tmg_model* m, int B, int h, int w, const float* x, const float* const* h_in,
                                     const float* const* c_in, const float* const* eps, const void* tape_v, const float* g_y,
                                     const float* g_log_det, const float* const* g_h_out, const float* const* g_c_out,
                                     float* const* g_h_in, float* const* g_c_in, float* grads, void* workspace,
                                     size_t workspace_bytes, uint32_t flags, void* stream) {
  if (!m) { set_error("null model"); return TMG_ERR_NULL; }
  Plan p;
  TMG_TRY(make_plan(*m, B, h, w, p, false));
  TMG_TRY(check_common(m, workspace, workspace_bytes, p));
  if (!x || !eps || !tape_v || !g_y || !g_log_det || !grads) { set_error("null argument"); return TMG_ERR_NULL; }
  if (workspace_bytes < tmg_reconstruct_backward_workspace_bytes(m, B, h, w)) { set_error("workspace too small for the backward pass"); return TMG_ERR_WORKSPACE; }
  const float* tape = (const float*)tape_v;
  // the recorded coupling-network intermediates are used only when this backward runs under the configuration the tape
  // was recorded with; otherwise every step recomputes them from its taped input
  tmg_model::TapeInfo tinfo;
  {
    std::lock_guard<std::mutex> lk(m->tape_mu);
    auto it = m->tapes.find(tape_v);
    if (it != m->tapes.end()) tinfo = it->second;
  }
  const bool tape_ok = (int)tinfo.emit.size() == m->cfg.n_levels && tinfo.sig[0] == B && tinfo.sig[1] == h &&
                       tinfo.sig[2] == w && tinfo.sig[3] == m->precision;
  const int L = p.L;
  Ctx c{*m, p, (float*)workspace, (cudaStream_t)stream};
  float* ws = c.ws;
  const RbExtra rx = rb_extra(*m, p);
  float* ex = (float*)((char*)workspace + align_up(p.total, 256));      // per-step scratch (BwdExtra layout)
  float* rb = ex + rx.e.total;                                           // chain buffers
  // conditioning maps and top prior parameters: recompute the encoder (batch statistics, running stats untouched)
  TMG_TRY(run_encoder(c, x, flags & TMG_FLAG_BN_TRAIN, 0.f));
  for (int l = 0; l < L; ++l)
    TMG_CUDA_OK(cudaMemsetAsync(rb + rx.gcond[l], 0, (size_t)B * p.Hl[l] * p.Wl[l] * m->cfg.cond_features * sizeof(float), c.st));
  float* Gc = rb + rx.ga;      // current gradient w.r.t. the flow state
  float* Gn = rb + rx.gb;
  {   // adjoint of the final CheckerSqueeze.reverse: g_y [B,out,H,W] -> [B,H/2,W/2,4*out]
    PermArgs pa{};
    pa.mode = PERM_SQUEEZE_NCHW_TO_NHWC; pa.src = g_y; pa.dst = Gc; pa.dst_cstride = m->levels[0].C; pa.dst_coff = 0;
    pa.B = B; pa.C = m->cfg.out_features; pa.H = p.H; pa.W = p.W;
    TMG_TRY(launch_permute(pa, c.st));
  }
  const int prec = m->precision;
  for (int l = 0; l < L; ++l) {
    const LevelW& lv = m->levels[l];
    const int Hl = p.Hl[l], Wl = p.Wl[l], HW = Hl * Wl, C = lv.C, n = (int)lv.steps.size();
    const BwdExtra e = bwd_extra(*m, l, B, Hl, Wl);
    for (int s = 0; s < n; ++s) {
      const StepW& st = lv.steps[s];
      StepBwdIO io{};
      io.Y = tape + tape_off(*m, p, l, s); io.COND = ws + p.cond[l]; io.GO = Gc; io.g_ld = g_log_det;
      if (tape_ok && tinfo.emit[l][s]) { io.D_tape = tape + tape_off_d(*m, p, l, s); io.HR_tape = tape + tape_off_h(*m, p, l, s); }
      io.GY = Gn; io.GC = rb + rx.gcond[l]; io.grads = grads;
      io.defer_lu = true;
      if (st.kind == STEP_LSTM) {
        io.h_prev = h_in ? h_in[l] : nullptr; io.c_prev = c_in ? c_in[l] : nullptr;
        io.g_hn = g_h_out ? g_h_out[l] : nullptr; io.g_cn = g_c_out ? g_c_out[l] : nullptr;
        io.g_hprev = (g_h_in && io.h_prev) ? g_h_in[l] : nullptr; io.g_cprev = (g_c_in && io.c_prev) ? g_c_in[l] : nullptr;
      }
      TMG_TRY(step_backward(c, l, st, B, Hl, Wl, io, ex, e));
      float* t = Gc; Gc = Gn; Gn = t;
    }
    // Split.reverse backward (flowUtils.py:316-335): Gc = gradient w.r.t. cat(z1, z2)
    const float* Ysplit = tape + tape_off(*m, p, l, n - 1);
    m->precision = TMG_PREC_FP32;
    int rc = run_split_prior(c, l, B, Hl, Wl, Ysplit);
    m->precision = prec;
    TMG_TRY(rc);
    GaussBwdArgs ga{};
    ga.prm = ws + p.hr; ga.prm_cstride = C;
    ga.g_val = Gc; ga.gv_cstride = C; ga.gv_coff = C / 2;
    ga.eps = eps[l]; ga.g_ld = g_log_det; ga.gain = c.Q() + lv.split_gain; ga.hardtanh = 1;
    ga.g_prm = ex + e.gz; ga.gp_cstride = C; ga.part = rb + rx.gpart;
    ga.B = B; ga.HW = HW; ga.n = C / 2;
    TMG_TRY(launch_gauss_bwd(ga, c.st));
    TMG_TRY(launch_reduce_cols(rb + rx.gpart, gauss_bwd_blocks(B, HW), 1, 0, 1, ex + e.tmp, 0, c.st));
    TMG_TRY(launch_scale_grad(ex + e.tmp, c.P() + lv.split_scale, grads + lv.split_scale, c.st));
    {
      const ConvSrc fs[1] = {{Ysplit, C, 0, C / 2, 0}};
      const BwdDest ds[1] = {{Gc, nullptr, C, 0, C / 2, 1}};
      TMG_TRY(conv_backward(c, B, Hl, Wl, lv.split, 1, fs, true, ex + e.gz, C, 0, ds, 1, grads, ex + e.wt, ex + e.wscr, ex + e.gscale));
    }
    if (l + 1 < L) {
      // z1 of this level is the un-squeezed output of the next one: adjoint = squeeze of the first C/2 channels
      PermArgs pa{};
      pa.mode = PERM_SQUEEZE_NHWC_TO_NHWC; pa.src = Gc; pa.src_cstride = C; pa.src_coff = 0;
      pa.dst = Gn; pa.dst_cstride = m->levels[l + 1].C; pa.dst_coff = 0;
      pa.B = B; pa.C = C / 2; pa.H = Hl; pa.W = Wl;
      TMG_TRY(launch_permute(pa, c.st));
      float* t = Gc; Gc = Gn; Gn = t;
    } else {
      // top latent z = cmean + exp(clamp(clog_std)) * eps[L]: gradient w.r.t. the encoder output [2*Cz]
      GaussBwdArgs gt{};
      gt.prm = ws + p.zout; gt.prm_cstride = 2 * m->Cz;
      gt.g_val = Gc; gt.gv_cstride = C; gt.gv_coff = 0;
      gt.eps = eps[L]; gt.g_ld = nullptr; gt.gain = nullptr; gt.hardtanh = 0;
      gt.g_prm = rb + rx.gzout; gt.gp_cstride = 2 * m->Cz; gt.part = nullptr;
      gt.B = B; gt.HW = HW; gt.n = m->Cz;
      TMG_TRY(launch_gauss_bwd(gt, c.st));
    }
  }
  // encoder parameters
  return run_encoder_backward(c, (flags & TMG_FLAG_BN_TRAIN) != 0, rb, rx, grads);
}

int tmg_reconstruct_backward(tmg_model* m, int B, int h, int w, const float* x, const float* const* h_in,
                             const float* const* c_in, const float* const* eps, const void* tape_v, const float* g_y,
                             const float* g_log_det, const float* const* g_h_out, const float* const* g_c_out,
                             float* const* g_h_in, float* const* g_c_in, float* grads, void* workspace,
                             size_t workspace_bytes, uint32_t flags, void* stream) {
  if (!m) { set_error("null model"); return TMG_ERR_NULL; }
  auto eager = [&](void* s_) {
    return reconstruct_backward_impl(m, B, h, w, x, h_in, c_in, eps, tape_v, g_y, g_log_det, g_h_out, g_c_out, g_h_in, g_c_in,
                                     grads, workspace, workspace_bytes, flags, s_);
  };
  // Opt-in (TMG_BWD_GRAPH=1): replay pays only when the caller's buffers keep their addresses from one optimizer step to
  // the next; torch's caching allocator moves the per-call tensors (tape, noise, gradients) often enough that captures
  // (~100 ms each) keep recurring in a short run -- measured: 193 -> 180 ms per step when the addresses repeat, slower
  // when they do not.
  static const bool use_graph = [] { const char* e = getenv("TMG_BWD_GRAPH"); return e && e[0] == '1'; }();
  if (!use_graph || !m->ready) { ++m->n_eager; return eager(stream); }
  // the launch sequence is a pure function of these values: same key -> same ~1 400 launches, replayed as one graph
  const int L = m->cfg.n_levels;
  std::vector<uint64_t> key;
  key.reserve(16 + 8 * L);
  auto put = [&](const void* p_) { key.push_back((uint64_t)(uintptr_t)p_); };
  key.push_back(((uint64_t)B << 40) ^ ((uint64_t)h << 20) ^ (uint64_t)w);
  key.push_back(((uint64_t)flags << 8) ^ (uint64_t)m->precision);
  key.push_back((uint64_t)workspace_bytes);
  put(m->params); put(m->packed); put(x); put(tape_v); put(g_y); put(g_log_det); put(grads); put(workspace);
  for (int l = 0; l <= L; ++l) put(eps ? eps[l] : nullptr);
  for (int l = 0; l < L; ++l) {
    put(h_in ? h_in[l] : nullptr); put(c_in ? c_in[l] : nullptr);
    put(g_h_out ? g_h_out[l] : nullptr); put(g_c_out ? g_c_out[l] : nullptr);
    put(g_h_in ? g_h_in[l] : nullptr); put(g_c_in ? g_c_in[l] : nullptr);
  }
  {   // which steps read their intermediates from the tape (host-side state of the last training forward)
    uint64_t hsh = 1469598103934665603ull;
    std::lock_guard<std::mutex> lk(m->tape_mu);
    auto it = m->tapes.find(tape_v);
    if (it != m->tapes.end()) {
      for (int q = 0; q < 4; ++q) hsh = (hsh ^ (uint64_t)(uint32_t)it->second.sig[q]) * 1099511628211ull;
      for (const auto& v : it->second.emit) for (char ch : v) hsh = (hsh ^ (uint64_t)(unsigned char)ch) * 1099511628211ull;
    }
    key.push_back(hsh);
  }
  if (m->bwd_graphs.size() > 1024) {       // keys that never repeated (unstable buffer addresses): forget them
    for (auto it = m->bwd_graphs.begin(); it != m->bwd_graphs.end();)
      it = it->second.exec ? std::next(it) : m->bwd_graphs.erase(it);
  }
  tmg_model::BwdGraph& g = m->bwd_graphs[key];
  cudaStream_t st = (cudaStream_t)stream;
  auto replay = [&]() -> int {
    TMG_CUDA_OK(cudaEventRecord(m->gev_in, st));
    TMG_CUDA_OK(cudaStreamWaitEvent(m->gstream, m->gev_in, 0));
    TMG_CUDA_OK(cudaGraphLaunch(g.exec, m->gstream));
    TMG_CUDA_OK(cudaEventRecord(m->gev_out, m->gstream));
    TMG_CUDA_OK(cudaStreamWaitEvent(st, m->gev_out, 0));
    tmg::g_launches += g.kernels;             // the kernels the graph launches (tmg_launch_count reports kernels)
    ++m->n_replays;
    return TMG_OK;
  };
  if (g.exec) return replay();
  if (g.seen++ == 0 || g.seen < 0 || m->n_graphs >= 64) { ++m->n_eager; return eager(stream); }      // first sight (or capture gave up): eager
  if (!m->gstream) {
    TMG_CUDA_OK(cudaStreamCreateWithFlags(&m->gstream, cudaStreamNonBlocking));
    TMG_CUDA_OK(cudaEventCreateWithFlags(&m->gev_in, cudaEventDisableTiming));
    TMG_CUDA_OK(cudaEventCreateWithFlags(&m->gev_out, cudaEventDisableTiming));
  }
  if (cudaStreamBeginCapture(m->gstream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    g.seen = -1000000;
    return eager(stream);
  }
  const int64_t launches_before = tmg::g_launches.load();
  const int rc = eager((void*)m->gstream);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(m->gstream, &graph);
  g.kernels = tmg::g_launches.load() - launches_before;
  tmg::g_launches.store(launches_before);          // nothing ran yet: the replay adds them
  if (rc != TMG_OK || ce != cudaSuccess || !graph) {
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    g.seen = -1000000;
    return rc != TMG_OK ? rc : eager(stream);
  }
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess || !exec) { cudaGetLastError(); g.seen = -1000000; return eager(stream); }
  g.exec = exec;
  ++m->n_graphs;
  return replay();
}

// ===================================================================== time-batched BPTT block
// TrainFlow.trainParallel runs `tback` sample() calls in a row and back-propagates through all of them
// (nn/trainFlowParallel.py:248-277).  Inside a level only the LSTM step couples the time steps (its ConvLSTM state);
// everything else of a time step depends on that time step's own input, noise and LSTM output.  So a BPTT block is
// restructured level by level: the LSTM step of the level runs T times in sequence on batch B, then the level's n-1 plain
// steps, its split prior and the squeeze run ONCE on batch T*B (time-major: slice t = samples [t*B, (t+1)*B)).  Same
// arithmetic per sample, 1/T of the launches for everything but the LSTM steps and the encoder (whose BatchNorm batch
// statistics are per sample() call in the reference, so it still runs per time step).
static Plan slice_plan(const tmg_model& m, const Plan& p, int t, int B) {
  Plan q = p;
  const tmg_config& g = m.cfg;
  const size_t tb = (size_t)t * B;
  q.B = B; q.Bx = B;
  q.xn += tb * p.h * p.w * g.in_features;
  q.e0 += tb * p.h * p.w * (g.init_features / 2);
  for (int l = 0; l < p.L; ++l) {
    q.db[l] += tb * p.eh[l] * p.ew[l] * m.levels[l].nf_out;
    q.cond[l] += tb * p.Hl[l] * p.Wl[l] * g.cond_features;
  }
  q.zout += tb * p.Hl[p.L - 1] * p.Wl[p.L - 1] * 2 * m.Cz;
  return q;      // cc / zo_pre / bn_* are scratch shared by the slices
}
// tape of a block: the per-step tape for batch T*B, then the ConvLSTM states after every time step [L][2][T*B,HW_l,R]
static size_t bptt_state_off(const tmg_model& m, const Plan& p, int l, int which) {
  size_t off = tape_floats(m, p);
  for (int q = 0; q < l; ++q) off += (size_t)2 * p.B * p.Hl[q] * p.Wl[q] * m.cfg.rec_features;
  return off + (size_t)which * p.B * p.Hl[l] * p.Wl[l] * m.cfg.rec_features;
}
static size_t bptt_tape_floats(const tmg_model& m, const Plan& p) { return bptt_state_off(m, p, p.L, 0); }

size_t tmg_bptt_tape_bytes(const tmg_model* m, int T, int B, int h, int w) {
  if (!m || T < 1) return 0;
  Plan p;
  if (make_plan(*m, T * B, h, w, p) != TMG_OK) return 0;
  return bptt_tape_floats(*m, p) * sizeof(float);
}
size_t tmg_bptt_workspace_bytes(const tmg_model* m, int T, int B, int h, int w) {
  if (!m || T < 1) return 0;
  Plan p;
  if (make_plan(*m, T * B, h, w, p) != TMG_OK) return 0;
  size_t mxs = 0;                                    // gradients w.r.t. the carried LSTM states: two (h, c) pairs
  for (int l = 0; l < p.L; ++l) mxs = std::max(mxs, (size_t)B * p.Hl[l] * p.Wl[l] * m->cfg.rec_features);
  return tmg_reconstruct_backward_workspace_bytes(m, T * B, h, w) + 4 * mxs * sizeof(float) + 256;
}

int tmg_bptt_forward(tmg_model* m, int T, int B, int h, int w, const float* x, const float* const* h_in,
                     const float* const* c_in, const float* const* eps, float* y, float* log_det, float* const* h_out,
                     float* const* c_out, void* tape_v, size_t tape_bytes, void* workspace, size_t workspace_bytes,
                     uint32_t flags, void* stream) {
  if (!m) { set_error("null model"); return TMG_ERR_NULL; }
  if (T < 1) { set_error("bad block length"); return TMG_ERR_BAD_SHAPE; }
  if (flags & TMG_FLAG_SHARED_X) { set_error("TMG_FLAG_SHARED_X is an inference option"); return TMG_ERR_BAD_CONFIG; }
  Plan p;
  const int BT = T * B;
  TMG_TRY(make_plan(*m, BT, h, w, p, false));
  TMG_TRY(check_common(m, workspace, workspace_bytes, p));
  if (!x || !eps || !y || !log_det || !h_out || !c_out || !tape_v) { set_error("null argument"); return TMG_ERR_NULL; }
  if (tape_bytes < bptt_tape_floats(*m, p) * sizeof(float)) { set_error("tape too small"); return TMG_ERR_WORKSPACE; }
  const int L = p.L, R = m->cfg.rec_features;
  for (int l = 0; l <= L; ++l) if (!eps[l]) { set_error("eps[%d] is null", l); return TMG_ERR_NULL; }
  float* tape = (float*)tape_v;
  Ctx c{*m, p, (float*)workspace, (cudaStream_t)stream};
  float* ws = c.ws;
  const int ldstride = p.nslots * p.ctas;
  TMG_CUDA_OK(cudaMemsetAsync(ws + p.ldp, 0, (size_t)BT * ldstride * sizeof(float), c.st));
  // encoder: once per time step (BatchNorm batch statistics and running-stat updates per sample() call, as in the reference)
  for (int t = 0; t < T; ++t) {
    Ctx ct{*m, slice_plan(*m, p, t, B), ws, c.st};
    TMG_TRY(run_encoder(ct, x + (size_t)t * B * m->cfg.in_features * h * w, flags & TMG_FLAG_BN_TRAIN));
  }
  if (prec_f16(m->precision)) TMG_TRY(run_hoist(c));
  tmg_model::TapeInfo tinfo;
  tinfo.emit.resize(L);
  for (int l = 0; l < L; ++l) tinfo.emit[l].assign(m->levels[l].steps.size(), 0);
  tinfo.sig[0] = BT; tinfo.sig[1] = h; tinfo.sig[2] = w; tinfo.sig[3] = m->precision;
  { std::lock_guard<std::mutex> lk(m->tape_mu); m->tapes.erase(tape); }
  {
    const LevelW& lv = m->levels[L - 1];
    GaussArgs ga{};
    ga.prm = ws + p.zout; ga.prm_cstride = 2 * m->Cz;
    ga.val = ws + p.y[L - 1]; ga.val_cstride = lv.C; ga.val_coff = 0;
    ga.eps_in = eps[L]; ga.reverse = 1;
    ga.B = BT; ga.HW = p.Hl[L - 1] * p.Wl[L - 1]; ga.n = m->Cz;
    ga.ld_part = nullptr; ga.ld_stride = ldstride;
    TMG_TRY(launch_gaussian(ga, c.st));
  }
  int slot = 0;
  for (int l = L - 1; l >= 0; --l) {
    const LevelW& lv = m->levels[l];
    const int Hl = p.Hl[l], Wl = p.Wl[l], HW = Hl * Wl, n = (int)lv.steps.size();
    float* Y = ws + p.y[l];
    float* Y2 = ws + p.y2[l];
    TMG_TRY(run_split_prior(c, l, BT, Hl, Wl, Y));
    GaussArgs ga{};
    ga.prm = ws + p.hr; ga.prm_cstride = lv.C;
    ga.val = Y; ga.val_cstride = lv.C; ga.val_coff = lv.C / 2;
    ga.eps_in = eps[l]; ga.reverse = 1; ga.B = BT; ga.HW = HW; ga.n = lv.C / 2;
    ga.ld_part = ws + p.ldp + (size_t)(slot++) * p.ctas; ga.ld_stride = ldstride;
    TMG_TRY(launch_gaussian(ga, c.st));
    for (int s = n - 1; s >= 0; --s) {
      const StepW& st = lv.steps[s];
      TMG_CUDA_OK(cudaMemcpyAsync(tape + tape_off(*m, p, l, s), Y, (size_t)BT * HW * lv.C * sizeof(float), cudaMemcpyDeviceToDevice, c.st));
      float* ldp = ws + p.ldp + (size_t)(slot++) * p.ctas;
      if (st.kind == STEP_LSTM) {
        // the only step that couples the time steps: T launches sequences on batch B, states carried through the tape
        float* hs = tape + bptt_state_off(*m, p, l, 0);
        float* cs = tape + bptt_state_off(*m, p, l, 1);
        const size_t sst = (size_t)B * HW * R, yst = (size_t)B * HW * lv.C, cst = (size_t)B * HW * m->cfg.cond_features;
        bool swapped = false;                            // fused kernels write the other buffer, the fp32 path works in place
        for (int t = 0; t < T; ++t) {
          float* Yt = Y + t * yst; float* Y2t = Y2 + t * yst;
          TMG_TRY(run_step(c, l, st, &st, true, B, Hl, Wl, Yt, Y2t, ws + p.cond[l] + t * cst,
                           t ? hs + (t - 1) * sst : (h_in ? h_in[l] : nullptr), t ? cs + (t - 1) * sst : (c_in ? c_in[l] : nullptr),
                           hs + t * sst, cs + t * sst, ldp + (size_t)t * B * ldstride));
          const bool sw = Yt == Y2 + t * yst;
          if (t && sw != swapped) { set_error("bptt: inconsistent buffer use of the LSTM step"); return TMG_ERR_UNSUPPORTED; }
          swapped = sw;
        }
        if (swapped) { float* tmp = Y; Y = Y2; Y2 = tmp; }
        TMG_CUDA_OK(cudaMemcpyAsync(h_out[l], hs + (size_t)(T - 1) * sst, sst * sizeof(float), cudaMemcpyDeviceToDevice, c.st));
        TMG_CUDA_OK(cudaMemcpyAsync(c_out[l], cs + (size_t)(T - 1) * sst, sst * sizeof(float), cudaMemcpyDeviceToDevice, c.st));
      } else {
        c.emit_d = tape + tape_off_d(*m, p, l, s); c.emit_h = tape + tape_off_h(*m, p, l, s); c.emitted = false;
        TMG_TRY(run_step(c, l, st, &st, true, BT, Hl, Wl, Y, Y2, ws + p.cond[l], nullptr, nullptr, nullptr, nullptr, ldp));
        tinfo.emit[l][s] = c.emitted ? 1 : 0;
        c.emit_d = c.emit_h = nullptr;
      }
    }
    PermArgs pa{};
    pa.src = Y; pa.src_cstride = lv.C; pa.src_coff = 0;
    pa.B = BT; pa.C = lv.C / 4; pa.H = 2 * Hl; pa.W = 2 * Wl;
    if (l > 0) {
      pa.mode = PERM_UNSQUEEZE_NHWC_TO_NHWC;
      pa.dst = ws + p.y[l - 1]; pa.dst_cstride = m->levels[l - 1].C; pa.dst_coff = 0;
    } else {
      pa.mode = PERM_UNSQUEEZE_NHWC_TO_NCHW;
      pa.dst = y;
    }
    TMG_TRY(launch_permute(pa, c.st));
  }
  {
    std::lock_guard<std::mutex> lk(m->tape_mu);
    tinfo.serial = ++m->tape_serial;
    m->tapes[tape] = std::move(tinfo);
  }
  return finish_logdet(c, log_det);
}

int tmg_bptt_backward(tmg_model* m, int T, int B, int h, int w, const float* x, const float* const* h_in,
                      const float* const* c_in, const float* const* eps, const void* tape_v, const float* g_y,
                      const float* g_log_det, const float* const* g_h_out, const float* const* g_c_out,
                      float* const* g_h_in, float* const* g_c_in, float* grads, void* workspace, size_t workspace_bytes,
                      uint32_t flags, void* stream) {
  if (!m) { set_error("null model"); return TMG_ERR_NULL; }
  if (T < 1) { set_error("bad block length"); return TMG_ERR_BAD_SHAPE; }
  Plan p;
  const int BT = T * B;
  TMG_TRY(make_plan(*m, BT, h, w, p, false));
  TMG_TRY(check_common(m, workspace, workspace_bytes, p));
  if (!x || !eps || !tape_v || !g_y || !g_log_det || !grads) { set_error("null argument"); return TMG_ERR_NULL; }
  if (workspace_bytes < tmg_bptt_workspace_bytes(m, T, B, h, w)) { set_error("workspace too small for the backward pass"); return TMG_ERR_WORKSPACE; }
  const float* tape = (const float*)tape_v;
  tmg_model::TapeInfo tinfo;
  {
    std::lock_guard<std::mutex> lk(m->tape_mu);
    auto it = m->tapes.find(tape_v);
    if (it != m->tapes.end()) tinfo = it->second;
  }
  const bool tape_ok = (int)tinfo.emit.size() == m->cfg.n_levels && tinfo.sig[0] == BT && tinfo.sig[1] == h &&
                       tinfo.sig[2] == w && tinfo.sig[3] == m->precision;
  const int L = p.L, R = m->cfg.rec_features, cf = m->cfg.cond_features;
  Ctx c{*m, p, (float*)workspace, (cudaStream_t)stream};
  float* ws = c.ws;
  const RbExtra rx = rb_extra(*m, p);
  float* ex = (float*)((char*)workspace + align_up(p.total, 256));
  float* rb = ex + rx.e.total;
  float* carry = rb + rx.total;                      // gradients w.r.t. the carried LSTM states: two (h, c) pairs, ping-pong
  size_t mxs = 0;
  for (int l = 0; l < L; ++l) mxs = std::max(mxs, (size_t)B * p.Hl[l] * p.Wl[l] * R);
  if ((size_t)((char*)(carry + 4 * mxs) - (char*)workspace) > workspace_bytes) { set_error("workspace too small for the state gradients"); return TMG_ERR_WORKSPACE; }
  const bool bn_train = (flags & TMG_FLAG_BN_TRAIN) != 0;
  for (int t = 0; t < T; ++t) {                       // conditioning maps and top prior parameters of every time step
    Ctx ct{*m, slice_plan(*m, p, t, B), ws, c.st};
    TMG_TRY(run_encoder(ct, x + (size_t)t * B * m->cfg.in_features * h * w, bn_train, 0.f));
  }
  for (int l = 0; l < L; ++l)
    TMG_CUDA_OK(cudaMemsetAsync(rb + rx.gcond[l], 0, (size_t)BT * p.Hl[l] * p.Wl[l] * cf * sizeof(float), c.st));
  float* Gc = rb + rx.ga;
  float* Gn = rb + rx.gb;
  {
    PermArgs pa{};
    pa.mode = PERM_SQUEEZE_NCHW_TO_NHWC; pa.src = g_y; pa.dst = Gc; pa.dst_cstride = m->levels[0].C; pa.dst_coff = 0;
    pa.B = BT; pa.C = m->cfg.out_features; pa.H = p.H; pa.W = p.W;
    TMG_TRY(launch_permute(pa, c.st));
  }
  const int prec = m->precision;
  for (int l = 0; l < L; ++l) {
    const LevelW& lv = m->levels[l];
    const int Hl = p.Hl[l], Wl = p.Wl[l], HW = Hl * Wl, C = lv.C, n = (int)lv.steps.size();
    const BwdExtra e = bwd_extra(*m, l, BT, Hl, Wl);
    for (int s = 0; s < n; ++s) {
      const StepW& st = lv.steps[s];
      if (st.kind != STEP_LSTM) {
        StepBwdIO io{};
        io.Y = tape + tape_off(*m, p, l, s); io.COND = ws + p.cond[l]; io.GO = Gc; io.g_ld = g_log_det;
        if (tape_ok && tinfo.emit[l][s]) { io.D_tape = tape + tape_off_d(*m, p, l, s); io.HR_tape = tape + tape_off_h(*m, p, l, s); }
        io.GY = Gn; io.GC = rb + rx.gcond[l]; io.grads = grads; io.defer_lu = true;
        TMG_TRY(step_backward(c, l, st, BT, Hl, Wl, io, ex, e));
      } else {
        const float* hs = tape + bptt_state_off(*m, p, l, 0);
        const float* cs = tape + bptt_state_off(*m, p, l, 1);
        const size_t sst = (size_t)B * HW * R, yst = (size_t)B * HW * C, cst = (size_t)B * HW * cf;
        for (int t = T - 1; t >= 0; --t) {
          StepBwdIO io{};
          io.Y = tape + tape_off(*m, p, l, s) + t * yst; io.COND = ws + p.cond[l] + t * cst;
          io.GO = Gc + t * yst; io.g_ld = g_log_det + (size_t)t * B;
          io.GY = Gn + t * yst; io.GC = rb + rx.gcond[l] + t * cst; io.grads = grads; io.defer_lu = true;
          io.h_prev = t ? hs + (t - 1) * sst : (h_in ? h_in[l] : nullptr);
          io.c_prev = t ? cs + (t - 1) * sst : (c_in ? c_in[l] : nullptr);
          float* ch_in = carry + (size_t)((t + 1) & 1) * 2 * mxs;        // written by time step t + 1
          float* ch_out = carry + (size_t)(t & 1) * 2 * mxs;
          io.g_hn = t == T - 1 ? (g_h_out ? g_h_out[l] : nullptr) : ch_in;
          io.g_cn = t == T - 1 ? (g_c_out ? g_c_out[l] : nullptr) : ch_in + mxs;
          io.g_hprev = t ? ch_out : ((g_h_in && io.h_prev) ? g_h_in[l] : nullptr);
          io.g_cprev = t ? ch_out + mxs : ((g_c_in && io.c_prev) ? g_c_in[l] : nullptr);
          TMG_TRY(step_backward(c, l, st, B, Hl, Wl, io, ex, e));
        }
      }
      float* tsw = Gc; Gc = Gn; Gn = tsw;
    }
    const float* Ysplit = tape + tape_off(*m, p, l, n - 1);
    m->precision = TMG_PREC_FP32;
    int rc = run_split_prior(c, l, BT, Hl, Wl, Ysplit);
    m->precision = prec;
    TMG_TRY(rc);
    GaussBwdArgs ga{};
    ga.prm = ws + p.hr; ga.prm_cstride = C;
    ga.g_val = Gc; ga.gv_cstride = C; ga.gv_coff = C / 2;
    ga.eps = eps[l]; ga.g_ld = g_log_det; ga.gain = c.Q() + lv.split_gain; ga.hardtanh = 1;
    ga.g_prm = ex + e.gz; ga.gp_cstride = C; ga.part = rb + rx.gpart;
    ga.B = BT; ga.HW = HW; ga.n = C / 2;
    TMG_TRY(launch_gauss_bwd(ga, c.st));
    TMG_TRY(launch_reduce_cols(rb + rx.gpart, gauss_bwd_blocks(BT, HW), 1, 0, 1, ex + e.tmp, 0, c.st));
    TMG_TRY(launch_scale_grad(ex + e.tmp, c.P() + lv.split_scale, grads + lv.split_scale, c.st));
    {
      const ConvSrc fs[1] = {{Ysplit, C, 0, C / 2, 0}};
      const BwdDest ds[1] = {{Gc, nullptr, C, 0, C / 2, 1}};
      TMG_TRY(conv_backward(c, BT, Hl, Wl, lv.split, 1, fs, true, ex + e.gz, C, 0, ds, 1, grads, ex + e.wt, ex + e.wscr, ex + e.gscale));
    }
    if (l + 1 < L) {
      PermArgs pa{};
      pa.mode = PERM_SQUEEZE_NHWC_TO_NHWC; pa.src = Gc; pa.src_cstride = C; pa.src_coff = 0;
      pa.dst = Gn; pa.dst_cstride = m->levels[l + 1].C; pa.dst_coff = 0;
      pa.B = BT; pa.C = C / 2; pa.H = Hl; pa.W = Wl;
      TMG_TRY(launch_permute(pa, c.st));
      float* tsw = Gc; Gc = Gn; Gn = tsw;
    } else {
      GaussBwdArgs gt{};
      gt.prm = ws + p.zout; gt.prm_cstride = 2 * m->Cz;
      gt.g_val = Gc; gt.gv_cstride = C; gt.gv_coff = 0;
      gt.eps = eps[L]; gt.g_ld = nullptr; gt.gain = nullptr; gt.hardtanh = 0;
      gt.g_prm = rb + rx.gzout; gt.gp_cstride = 2 * m->Cz; gt.part = nullptr;
      gt.B = BT; gt.HW = HW; gt.n = m->Cz;
      TMG_TRY(launch_gauss_bwd(gt, c.st));
    }
  }
  // encoder parameters: per time step (batch statistics of that sample() call)
  for (int t = 0; t < T; ++t) {
    Ctx ct{*m, slice_plan(*m, p, t, B), ws, c.st};
    RbExtra rxt = rx;
    const size_t tb = (size_t)t * B;
    for (int l = 0; l < L; ++l) rxt.gcond[l] += tb * p.Hl[l] * p.Wl[l] * cf;
    rxt.gzout += tb * p.Hl[L - 1] * p.Wl[L - 1] * 2 * m->Cz;
    TMG_TRY(run_encoder_backward(ct, bn_train, rb, rxt, grads));
  }
  return TMG_OK;
}

// how the backward calls of this model ran so far: captured graphs, graph replays, eager launch sequences
int tmg_backward_graph_stats(const tmg_model* m, int64_t* graphs, int64_t* replays, int64_t* eager) {
  if (!m) { set_error("null model"); return TMG_ERR_NULL; }
  if (graphs) *graphs = m->n_graphs;
  if (replays) *replays = m->n_replays;
  if (eager) *eager = m->n_eager;
  return TMG_OK;
}

// Finishes the parameter gradients tmg_reconstruct_backward defers: the LU-parameterised 1x1 convolutions
// (glowConv.py:151-174: l, u, log_s) and the log-det terms of log_s / ActNorm weights, for all flow steps in one launch.
int tmg_backward_finalize(tmg_model* m, float* grads, void* stream) {
  if (!m || !grads) { set_error("null argument"); return TMG_ERR_NULL; }
  if (!m->ready || !m->lu_tab_dev) { set_error("tmg_model_refresh() has not been called"); return TMG_ERR_NOT_READY; }
  return launch_lu_bwd_batched(m->lu_tab_dev, (int)m->lu_tab.size(), m->cmax, m->params, grads, (cudaStream_t)stream);
}

size_t tmg_flow_step_backward_workspace_bytes(tmg_model* m, int level, int B, int Hl, int Wl) {
  Plan p;
  if (!m || op_plan(m, level, B, Hl, Wl, p) != TMG_OK) return 0;
  return p.total + bwd_extra(*m, level, B, Hl, Wl).total * sizeof(float) + 256;
}

int tmg_flow_step_backward(tmg_model* m, int level, int step, int B, int Hl, int Wl, const float* x, const float* cond,
                           const float* h_in, const float* c_in, const float* g_out, const float* g_logdet,
                           const float* g_h_out, const float* g_c_out, float* g_x, float* g_cond, float* g_h_in,
                           float* g_c_in, float* grads, void* workspace, size_t workspace_bytes, void* stream) {
  Plan p;
  TMG_TRY(op_plan(m, level, B, Hl, Wl, p));
  TMG_TRY(check_common(m, workspace, workspace_bytes, p));
  if (!x || !cond || !g_out || !g_logdet || !g_x || !g_cond || !grads) { set_error("null argument"); return TMG_ERR_NULL; }
  const LevelW& lv = m->levels[level];
  if (step < 1 || step > (int)lv.steps.size()) { set_error("bad step %d", step); return TMG_ERR_BAD_SHAPE; }
  const StepW& st = lv.steps[step - 1];
  const BwdExtra e = bwd_extra(*m, level, B, Hl, Wl);
  if (workspace_bytes < align_up(p.total, 256) + e.total * sizeof(float)) { set_error("workspace too small for the backward pass"); return TMG_ERR_WORKSPACE; }
  Ctx c{*m, p, (float*)workspace, (cudaStream_t)stream};
  float* ws = c.ws;
  float* ex = (float*)((char*)workspace + align_up(p.total, 256));
  const int HW = Hl * Wl, C = lv.C, cf = m->cfg.cond_features;
  float* Y = ws + p.scratch_in;
  float* CN = ws + p.scratch_cond;
  float *GO = ex + e.go, *GY = ex + e.gy, *GC = ex + e.gcond;
  PermArgs pa{};
  pa.mode = PERM_NCHW_TO_NHWC; pa.B = B; pa.H = Hl; pa.W = Wl;
  pa.src = x; pa.dst = Y; pa.C = C; pa.dst_cstride = C; TMG_TRY(launch_permute(pa, c.st));
  pa.src = g_out; pa.dst = GO; TMG_TRY(launch_permute(pa, c.st));
  pa.src = cond; pa.dst = CN; pa.C = cf; pa.dst_cstride = cf; TMG_TRY(launch_permute(pa, c.st));
  TMG_CUDA_OK(cudaMemsetAsync(GC, 0, (size_t)B * HW * cf * sizeof(float), c.st));
  StepBwdIO io{};
  io.Y = Y; io.COND = CN; io.GO = GO; io.g_ld = g_logdet; io.GY = GY; io.GC = GC;
  io.h_prev = h_in; io.c_prev = c_in; io.g_hn = g_h_out; io.g_cn = g_c_out; io.g_hprev = g_h_in; io.g_cprev = g_c_in;
  io.grads = grads;
  TMG_TRY(step_backward(c, level, st, B, Hl, Wl, io, ex, e));
  PermArgs pb{};
  pb.mode = PERM_NHWC_TO_NCHW; pb.B = B; pb.H = Hl; pb.W = Wl;
  pb.src = GY; pb.dst = g_x; pb.C = C; pb.src_cstride = C; TMG_TRY(launch_permute(pb, c.st));
  pb.src = GC; pb.dst = g_cond; pb.C = cf; pb.src_cstride = cf; TMG_TRY(launch_permute(pb, c.st));
  return TMG_OK;
}

static int split_common(tmg_model* m, int level, int B, int Hl, int Wl, const float* zin, int cin_ch,
                        const float* eps_in, float* eps_out, float* z1_out, float* z_out, float* logp,
                        void* workspace, size_t workspace_bytes, void* stream, bool reverse) {
  Plan p;
  TMG_TRY(op_plan(m, level, B, Hl, Wl, p));
  TMG_TRY(check_common(m, workspace, workspace_bytes, p));
  const LevelW& lv = m->levels[level];
  Ctx c{*m, p, (float*)workspace, (cudaStream_t)stream};
  float* ws = c.ws;
  const int HW = Hl * Wl, C = lv.C;
  float* Y = ws + p.scratch_in;
  const int ldstride = p.nslots * p.ctas;
  TMG_CUDA_OK(cudaMemsetAsync(ws + p.ldp, 0, (size_t)B * ldstride * sizeof(float), c.st));
  PermArgs pa{};
  pa.src = zin; pa.dst = Y; pa.mode = PERM_NCHW_TO_NHWC; pa.B = B; pa.C = cin_ch; pa.H = Hl; pa.W = Wl; pa.dst_cstride = C;
  TMG_TRY(launch_permute(pa, c.st));
  TMG_TRY(run_split_prior(c, level, B, Hl, Wl, Y));
  GaussArgs ga{};
  ga.prm = ws + p.hr; ga.prm_cstride = C;
  ga.val = Y; ga.val_cstride = C; ga.val_coff = C / 2;
  ga.eps_in = eps_in; ga.eps_out = eps_out; ga.reverse = reverse ? 1 : 0;
  ga.B = B; ga.HW = HW; ga.n = C / 2; ga.ld_part = ws + p.ldp; ga.ld_stride = ldstride;
  TMG_TRY(launch_gaussian(ga, c.st));
  PermArgs pb{};
  pb.src = Y; pb.mode = PERM_NHWC_TO_NCHW; pb.B = B; pb.H = Hl; pb.W = Wl; pb.src_cstride = C;
  if (reverse) { pb.dst = z_out; pb.C = C; } else { pb.dst = z1_out; pb.C = C / 2; }
  TMG_TRY(launch_permute(pb, c.st));
  LogdetArgs a{};
  a.ld_part = ws + p.ldp; a.ld_stride = ldstride; a.step_const = c.Q() + m->step_const_off;
  a.n_levels = 0; a.out = logp; a.B = B;
  return launch_logdet_reduce(a, c.st);
}

int tmg_split_forward(tmg_model* m, int level, int B, int Hl, int Wl, const float* z, float* z1, float* logp,
                      float* eps, void* workspace, size_t workspace_bytes, void* stream) {
  if (!m || !z || !z1 || !logp) { set_error("null argument"); return TMG_ERR_NULL; }
  if (level < 0 || level >= m->cfg.n_levels) { set_error("bad level"); return TMG_ERR_BAD_SHAPE; }
  return split_common(m, level, B, Hl, Wl, z, m->levels[level].C, nullptr, eps, z1, nullptr, logp,
                      workspace, workspace_bytes, stream, false);
}

int tmg_split_reverse(tmg_model* m, int level, int B, int Hl, int Wl, const float* z1, const float* eps, float* z,
                      float* logp, void* workspace, size_t workspace_bytes, void* stream) {
  if (!m || !z1 || !eps || !z || !logp) { set_error("null argument"); return TMG_ERR_NULL; }
  if (level < 0 || level >= m->cfg.n_levels) { set_error("bad level"); return TMG_ERR_BAD_SHAPE; }
  return split_common(m, level, B, Hl, Wl, z1, m->levels[level].C / 2, eps, nullptr, nullptr, z, logp,
                      workspace, workspace_bytes, stream, true);
}

}  // extern "C"
