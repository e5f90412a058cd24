// extern "C" boundary of libsrdiff_b200.so: error reporting and convolution plans.
#include <string.h>
#include <stddef.h>
#include <cuda.h>
#include <new>
#include <vector>
#include "epilogue.cuh"

namespace srd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int launch_conv_direct(const SrdConvDesc& d, const float* temb, cudaStream_t stream);

struct TcPlan;
int tc_supported(const SrdConvDesc& d, const char** why);
TcPlan* tc_plan_new();
void tc_plan_delete(TcPlan*);
int tc_plan_build(const SrdConvDesc& d, TcPlan* plan, int* err_flag);
int launch_conv_tc(const TcPlan& plan, const float* temb, cudaStream_t stream);
void tc_set_mode(int a_mode, int b_resident);
int tc_s2_halo();
struct TcChain;
struct TcChainLink {
  int src_layer[SRD_MAX_SRC], src_which[SRD_MAX_SRC];
  int res1_layer, res1_which;
  int temb_off;
};
TcChain* tc_chain_new();
void tc_chain_delete(TcChain*);
int tc_chain_build(const TcPlan* const* plans, const TcChainLink* links, int n, TcChain* chain, int single_segments);
int tc_chain_segments(const TcChain* c);
int tc_chain_max_layers();
int tc_same_segment(const TcPlan* x, const TcPlan* y);
int tc_seg_k();
int launch_chain(const TcChain& c, const float* temb_row, cudaStream_t stream);
void tc_set_debug(unsigned long long* p);

static int* g_err_flag_dev = nullptr;   // device-visible pointer to a host-mapped int
static int* g_err_flag_host = nullptr;

static int* watchdog_flag() {
  if (!g_err_flag_host) {
    if (cudaHostAlloc((void**)&g_err_flag_host, sizeof(int), cudaHostAllocMapped) != cudaSuccess) {
      g_err_flag_host = nullptr;
      cudaGetLastError();
      return nullptr;
    }
    *g_err_flag_host = 0;
    if (cudaHostGetDevicePointer((void**)&g_err_flag_dev, g_err_flag_host, 0) != cudaSuccess) {
      g_err_flag_dev = nullptr;
      cudaGetLastError();
    }
  }
  return g_err_flag_dev;
}

}  // namespace srd

using namespace srd;

struct SrdConvPlan {
  SrdConvDesc desc;
  int kernel;
  srd::TcPlan* tc;
};

static int validate(const SrdConvDesc& d) {
  SRD_CHECK(d.dtype == SRD_F32 || d.dtype == SRD_BF16, "conv: bad dtype %d", d.dtype);
  SRD_CHECK(d.n_src >= 1 && d.n_src <= SRD_MAX_SRC, "conv: n_src %d out of range", d.n_src);
  SRD_CHECK(d.B > 0 && d.H_in > 0 && d.W_in > 0 && d.N > 0, "conv: bad sizes");
  SRD_CHECK(d.stride == 1 || d.stride == 2, "conv: stride must be 1 or 2");
  SRD_CHECK(d.up == 1 || d.up == 2 || d.up == 4 || d.up == 8, "conv: up must be 1, 2, 4 or 8");
  SRD_CHECK(d.up == 1 || (d.stride == 1 && d.nky == 2 && d.nkx == 2), "conv: transposed phases are 2x2 taps, stride 1");
  SRD_CHECK(d.nky >= 1 && d.nky <= 3 && d.nkx >= 1 && d.nkx <= 3, "conv: taps out of range");
  SRD_CHECK(d.up > 1 || (d.ky0 >= 0 && d.ky0 + d.nky <= 3 && d.kx0 >= 0 && d.kx0 + d.nkx <= 3),
            "conv: tap window outside the 1-pixel ring");
  SRD_CHECK(d.stride == 1 || (d.H_in % 2 == 0 && d.W_in % 2 == 0), "conv: stride 2 needs even dims");
  SRD_CHECK(d.H_out == d.H_in * d.up / d.stride && d.W_out == d.W_in * d.up / d.stride,
            "conv: H_out/W_out inconsistent with H_in/W_in (%dx%d -> %dx%d)", d.H_in, d.W_in, d.H_out, d.W_out);
  SRD_CHECK(d.border != SRD_BORDER_REFLECT || (d.H_out >= 2 && d.W_out >= 2), "conv: reflect ring needs H,W >= 2");
  SRD_CHECK(d.bias != nullptr, "conv: bias is NULL");
  SRD_CHECK(d.out || d.out2 || d.planar, "conv: no destination");
  for (int i = 0; i < d.n_src; ++i) {
    SRD_CHECK(d.src[i].ptr != nullptr, "conv: src[%d].ptr is NULL", i);
    if (d.src[i].flags & SRD_SRC_WIN8) continue;      // 64 = 8 pixels x 8 channels, checked by the tc planner
    SRD_CHECK(d.src[i].cin > 0 && d.src[i].c_off >= 0 && d.src[i].c_off + d.src[i].cin <= d.src[i].ct,
              "conv: src[%d] channel window [%d,+%d) outside ct=%d", i, d.src[i].c_off, d.src[i].cin, d.src[i].ct);
  }
  SRD_CHECK(!d.out || (d.out_coff >= 0 && d.out_coff + d.N <= d.out_ct), "conv: out channel window outside out_ct");
  SRD_CHECK(!d.out2 || (d.out2_coff >= 0 && d.out2_coff + d.N <= d.out2_ct), "conv: out2 channel window outside out2_ct");
  SRD_CHECK(!d.res1 || (d.res1_coff + d.N <= d.res1_ct), "conv: res1 channel window outside res1_ct");
  SRD_CHECK(!d.res2 || (d.res2_coff + d.N <= d.res2_ct), "conv: res2 channel window outside res2_ct");
  SRD_CHECK(!d.planar || d.planar_c > 0, "conv: planar_c must be positive");
  SRD_CHECK(!d.head_w || (d.head_b && d.planar), "conv: fused head needs head_b and planar");
  SRD_CHECK(d.N2 == 0 || (d.N2 == d.N && d.bias2 && d.dst2 && d.dst2_coff >= 0 && d.dst2_coff + d.N2 <= d.dst2_ct),
            "conv: second output needs N2 == N, bias2 and a dst2 channel window inside dst2_ct");
  SRD_CHECK(!d.lin_src || (d.lin_w && d.lin_c >= 1 && d.lin_c <= 4 && d.lin_c <= d.lin_ct),
            "conv: linear residual needs lin_w and 1 <= lin_c <= min(4, lin_ct)");
  return 0;
}

extern "C" {

const char* srd_last_error(void) { return g_err; }

int srd_version(void) { return 100; }

int srd_abi_sizeof(int32_t which) {
  switch (which) {
    case 0: return (int)sizeof(SrdConvDesc);
    case 1: return (int)sizeof(SrdPSampleArgs);
    case 2: return (int)offsetof(SrdConvDesc, dst2);
    case 3: return (int)offsetof(SrdPSampleArgs, seed_ptr);
    case 4: return (int)sizeof(SrdGroupNormArgs);
    case 5: return (int)offsetof(SrdGroupNormArgs, stats);
    default: return -1;
  }
}

int srd_conv_tc_supported(const SrdConvDesc* desc) {
  if (!desc) return 0;
  return tc_supported(*desc, nullptr);
}

int srd_conv_plan_create(const SrdConvDesc* desc, SrdConvPlan** out_plan) {
  SRD_CHECK(desc && out_plan, "srd_conv_plan_create: NULL argument");
  if (validate(*desc)) return 1;
  SrdConvPlan* p = new (std::nothrow) SrdConvPlan;
  SRD_CHECK(p != nullptr, "srd_conv_plan_create: out of memory");
  p->desc = *desc;
  p->tc = nullptr;
  int k = desc->kernel;
  if (k == SRD_KERNEL_AUTO) k = tc_supported(*desc, nullptr) ? SRD_KERNEL_TC : SRD_KERNEL_DIRECT;
  if (k == SRD_KERNEL_TC) {
    p->tc = tc_plan_new();
    if (!p->tc || tc_plan_build(*desc, p->tc, watchdog_flag())) {
      if (p->tc) tc_plan_delete(p->tc);
      delete p;
      return 1;
    }
  } else {
    if (!desc->w_direct) { delete p; SRD_CHECK(false, "srd_conv_plan_create: direct kernel needs w_direct"); }
    if (desc->head_w) { delete p; SRD_CHECK(false, "srd_conv_plan_create: fused head needs the tc kernel"); }
    if (desc->lin_src) { delete p; SRD_CHECK(false, "srd_conv_plan_create: linear residual needs the tc kernel"); }
    if (desc->N2) { delete p; SRD_CHECK(false, "srd_conv_plan_create: the second output needs the tc kernel"); }
    for (int i = 0; i < desc->n_src; ++i)
      if (desc->src[i].flags & SRD_SRC_WIN8) { delete p; SRD_CHECK(false, "srd_conv_plan_create: SRD_SRC_WIN8 needs the tc kernel"); }
  }
  p->kernel = k;
  *out_plan = p;
  return 0;
}

int srd_conv_plan_destroy(SrdConvPlan* plan) {
  if (!plan) return 0;
  if (plan->tc) tc_plan_delete(plan->tc);
  delete plan;
  return 0;
}

int srd_conv_plan_kernel(const SrdConvPlan* plan) { return plan ? plan->kernel : -1; }

int srd_conv_run(const SrdConvPlan* plan, const float* temb, void* stream) {
  SRD_CHECK(plan != nullptr, "srd_conv_run: NULL plan");
  if (plan->kernel == SRD_KERNEL_TC) return launch_conv_tc(*plan->tc, temb, (cudaStream_t)stream);
  return launch_conv_direct(plan->desc, temb, (cudaStream_t)stream);
}

int srd_tc_set_mode(int32_t a_mode, int32_t b_resident) {
  SRD_CHECK(a_mode >= 0 && a_mode <= 127, "srd_tc_set_mode: a_mode must be in 0..127");
  tc_set_mode(a_mode, b_resident ? 1 : 0);
  return 0;
}

int srd_tc_s2_halo(void) { return tc_s2_halo(); }

int srd_tc_set_debug(void* buf) {
  tc_set_debug((unsigned long long*)buf);
  return 0;
}

int srd_chain_create(SrdConvPlan* const* plans, const SrdChainLink* links, int32_t n, int32_t flags, SrdChain** out) {
  SRD_CHECK(plans && links && out && n > 0, "srd_chain_create: bad arguments");
  static_assert(sizeof(SrdChainLink) == sizeof(srd::TcChainLink), "SrdChainLink layout");
  std::vector<const srd::TcPlan*> tcs((size_t)n);
  for (int i = 0; i < n; ++i) {
    SRD_CHECK(plans[i] && plans[i]->kernel == SRD_KERNEL_TC && plans[i]->tc,
              "srd_chain_create: layer %d is not a tcgen05 plan (only those can be chained)", i);
    tcs[(size_t)i] = plans[i]->tc;
  }
  srd::TcChain* c = tc_chain_new();
  SRD_CHECK(c != nullptr, "srd_chain_create: out of memory");
  if (tc_chain_build(tcs.data(), reinterpret_cast<const srd::TcChainLink*>(links), n, c, flags)) {
    tc_chain_delete(c);
    return 1;
  }
  *out = reinterpret_cast<SrdChain*>(c);
  return 0;
}

int srd_chain_run(const SrdChain* chain, const float* temb_row, void* stream) {
  SRD_CHECK(chain != nullptr, "srd_chain_run: NULL chain");
  return launch_chain(*reinterpret_cast<const srd::TcChain*>(chain), temb_row, (cudaStream_t)stream);
}

int srd_chain_destroy(SrdChain* chain) {
  if (chain) tc_chain_delete(reinterpret_cast<srd::TcChain*>(chain));
  return 0;
}

int srd_chain_max_layers(void) { return tc_chain_max_layers(); }

int srd_chain_fused_max_layers(void) { return tc_seg_k(); }

int srd_conv_plans_share_segment(const SrdConvPlan* x, const SrdConvPlan* y) {
  if (!x || !y || x->kernel != SRD_KERNEL_TC || y->kernel != SRD_KERNEL_TC) return 0;
  return tc_same_segment(x->tc, y->tc);
}

int srd_chain_segments(const SrdChain* chain) {
  return chain ? tc_chain_segments(reinterpret_cast<const srd::TcChain*>(chain)) : -1;
}

/* Non-zero after a tcgen05 pipeline watchdog fired (code = which wait timed out). */
int srd_watchdog_flag(void) { return g_err_flag_host ? *g_err_flag_host : 0; }

}  // extern "C"
