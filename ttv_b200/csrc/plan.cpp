// plan.cpp -- argument validation, stride/layout folder, kernel chooser (pure host code).  See plan.h.
#include "plan.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace ttvb {

// ---- dtypes -------------------------------------------------------------------------------------------------
int dtype_size(int dtype)
{
  switch (dtype) {
    case TTV_B200_F32:  return 4;
    case TTV_B200_F64:  return 8;
    case TTV_B200_C64:  return 8;
    case TTV_B200_C128: return 16;
    case TTV_B200_I32:  return 4;
    case TTV_B200_I64:  return 8;
    default: return 0;
  }
}
int dtype_is_complex(int dtype) { return dtype == TTV_B200_C64 || dtype == TTV_B200_C128; }

// ---- L0 helpers ---------------------------------------------------------------------------------------------
// valid shape: at least one mode, no zero extent                                   (reference detail/shape.h:30-34)
bool is_valid_shape(const uint64_t* n, uint64_t p)
{
  if (p == 0) return false;
  return std::none_of(n, n + p, [](uint64_t x) { return x == 0; });
}

// valid layout: a permutation of 1..p                                             (reference detail/layout.h:29-55)
bool is_valid_layout(const uint64_t* pi, uint64_t p)
{
  if (p == 0) return false;
  for (uint64_t r = 0; r < p; ++r) {
    if (pi[r] < 1 || pi[r] > p) return false;
    if (std::find(pi + r + 1, pi + p, pi[r]) != pi + p) return false;
  }
  return true;
}

// strides never decrease when walking the modes in layout order                  (reference detail/strides.h:76-101)
bool is_valid_strides(const uint64_t* pi, uint64_t p, const uint64_t* w)
{
  for (uint64_t r = 1; r < p; ++r)
    if (w[pi[r] - 1] < w[pi[r - 1] - 1]) return false;
  return true;
}

static bool rest_is_one(const uint64_t* n, uint64_t from, uint64_t p)
{
  return std::all_of(n + std::min(from, p), n + p, [](uint64_t x) { return x == 1; });
}

// Packed strides of (n, pi) -- except that scalar- and vector-shaped tensors get all-one strides, which is what the
// reference hands out (detail/strides.h:42-45 with the shape predicates of detail/shape.h:38-63).
int compute_strides(const uint64_t* n, const uint64_t* pi, uint64_t p, uint64_t* w)
{
  if (!is_valid_shape(n, p) || !is_valid_layout(pi, p)) return -1;
  std::fill(w, w + p, uint64_t{1});
  const bool scalar = rest_is_one(n, 0, p);
  const bool vector = (p == 1 && n[0] > 1) ||
                      (p >= 2 && (n[0] > 1 || n[1] > 1) && (n[0] == 1 || n[1] == 1) && rest_is_one(n, 2, p));
  if (scalar || vector) return 0;
  for (uint64_t r = 1; r < p; ++r) {
    const uint64_t prev = pi[r - 1] - 1;
    w[pi[r] - 1] = w[prev] * n[prev];
  }
  return 0;
}

// output shape = input shape without entry q                                     (reference detail/shape.h:103-123)
int output_shape(const uint64_t* na, uint64_t p, uint64_t q, uint64_t* nc)
{
  if (!is_valid_shape(na, p) || q < 1 || q > p) return -1;
  std::copy(na, na + (q - 1), nc);
  std::copy(na + q, na + p, nc + (q - 1));
  return 0;
}

// output layout = input layout without q, modes above q renumbered             (reference detail/layout.h:143-172)
int output_layout(const uint64_t* pia, uint64_t p, uint64_t q, uint64_t* pic)
{
  if (!is_valid_layout(pia, p) || q < 1 || q > p) return -1;
  uint64_t* out = pic;
  for (uint64_t r = 0; r < p; ++r)
    if (pia[r] != q) *out++ = pia[r] - (pia[r] > q ? 1 : 0);
  return 0;
}

// k-order layout (k, k-1, .., 1, k+1, .., p); k = 0 or k > p => last-order        (reference detail/layout.h:57-76)
int k_order_layout(uint64_t p, uint64_t k, uint64_t* pi)
{
  if (p == 0) return -1;
  const uint64_t m = (k == 0 || k > p) ? p : k;
  for (uint64_t r = 0; r < p; ++r) pi[r] = r < m ? m - r : r + 1;
  return 0;
}

// the reference's 8 cases                                                         (reference detail/cases.h:24-36)
int classify_case(uint64_t p, uint64_t q, const uint64_t* pia)
{
  if (p == 1) return 1;
  if (p == 2) return pia[0] == 1 ? (q == 1 ? 2 : 3) : (q == 1 ? 4 : 5);
  if (pia[0] == q) return 6;
  if (pia[p - 1] == q) return 7;
  return 8;
}

// ---- messages -----------------------------------------------------------------------------------------------
const char* status_message(int status)
{
  switch (status) {
    case TTV_B200_OK: return "ok";
    case TTV_B200_ERR_ORDER_ZERO:      return "Error in tlib::tensor_times_vector: input tensor order should be greater zero.";
    case TTV_B200_ERR_MODE:            return "Error in tlib::tensor_times_vector: contraction mode should be greater zero or less than or equal to p.";
    case TTV_B200_ERR_A_NULL:          return "Error in tlib::tensor_times_vector: pointer to input tensor A should not be zero.";
    case TTV_B200_ERR_B_NULL:          return "Error in tlib::tensor_times_vector: pointer to input vector B should not be zero.";
    case TTV_B200_ERR_C_NULL:          return "Error in tlib::tensor_times_vector: pointer to output tensor C should not be zero.";
    case TTV_B200_ERR_NA_NULL:         return "Error in tlib::tensor_times_vector: pointer to input tensor shape vector na should not be zero.";
    case TTV_B200_ERR_NB_NULL:         return "Error in tlib::tensor_times_vector: pointer to input vector shape vector nb should not be zero.";
    case TTV_B200_ERR_NC_NULL:         return "Error in tlib::tensor_times_vector: pointer to output tensor shape vector nc should not be zero.";
    case TTV_B200_ERR_WA_NULL:         return "Error in tlib::tensor_times_vector: pointer to input tensor stride vector wa should not be zero.";
    case TTV_B200_ERR_WC_NULL:         return "Error in tlib::tensor_times_vector: pointer to output tensor stride vector wc should not be zero.";
    case TTV_B200_ERR_PIA_NULL:        return "Error in tlib::tensor_times_vector: pointer to input tensor permutation vector pia should not be zero.";
    case TTV_B200_ERR_PIC_NULL:        return "Error in tlib::tensor_times_vector: pointer to output tensor permutation vector pic should not be zero.";
    case TTV_B200_ERR_EXTENT_MISMATCH: return "Error in tlib::tensor_times_vector: contraction dimension of A and B are not equal.";
    case TTV_B200_ERR_SHAPE_A:         return "Error in tlib::tensor_times_vector: shape vector of A is not valid.";
    case TTV_B200_ERR_SHAPE_C:         return "Error in tlib::tensor_times_vector: shape vector of C is not valid.";
    case TTV_B200_ERR_LAYOUT_A:        return "Error in tlib::tensor_times_vector: layout vector of A is not valid.";
    case TTV_B200_ERR_LAYOUT_C:        return "Error in tlib::tensor_times_vector: layout vector of C is not valid.";
    case TTV_B200_ERR_STRIDES_A:       return "Error in tlib::tensor_times_vector: stride vector of A is not valid.";
    case TTV_B200_ERR_STRIDES_C:       return "Error in tlib::tensor_times_vector: stride vector of C is not valid.";
    case TTV_B200_ERR_LAYOUT_BEGIN:    return "Error in tlib::detail::compute_inverse_pia_m: beginning of layout tuples of both tensors are not correct.";
    case TTV_B200_ERR_LAYOUT_END:      return "Error in tlib::detail::compute_inverse_pia_m: end of layout tuples of both tensors are not correct.";
    case TTV_B200_ERR_NOT_PACKED:      return "Error in ttv_b200: strides / shape of A and C do not describe the same free modes (or more than 8 unfoldable ones).";
    case TTV_B200_ERR_DTYPE:           return "Error in ttv_b200: unknown element type.";
    case TTV_B200_ERR_OPTS:            return "Error in ttv_b200: invalid options.";
    case TTV_B200_ERR_CUDA:            return "Error in ttv_b200: CUDA failure (no CPU fallback exists).";
    case TTV_B200_ERR_MIXED_POINTERS:  return "Error in ttv_b200: a, b and c must be all host or all device pointers.";
    default: return "Error in ttv_b200: unknown status.";
  }
}

// ---- validation + folding -----------------------------------------------------------------------------------
int validate_and_fold(uint64_t q, uint64_t p,
                      const void* a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,
                      const void* b, const uint64_t* nb,
                      const void* c, const uint64_t* nc, const uint64_t* wc, const uint64_t* pic,
                      uint32_t flags, View* view)
{
  // the checks of the low-level interface in its order (reference ttv.h:64-89)
  if (p == 0)                                 return TTV_B200_ERR_ORDER_ZERO;
  if (q == 0 || q > p)                        return TTV_B200_ERR_MODE;
  if (a == nullptr)                           return TTV_B200_ERR_A_NULL;
  if (b == nullptr)                           return TTV_B200_ERR_B_NULL;
  if (c == nullptr)                           return TTV_B200_ERR_C_NULL;
  if (na == nullptr)                          return TTV_B200_ERR_NA_NULL;
  if (nb == nullptr)                          return TTV_B200_ERR_NB_NULL;
  if (nc == nullptr)                          return TTV_B200_ERR_NC_NULL;
  if (wa == nullptr)                          return TTV_B200_ERR_WA_NULL;
  if (wc == nullptr)                          return TTV_B200_ERR_WC_NULL;
  if (pia == nullptr)                         return TTV_B200_ERR_PIA_NULL;
  if (pic == nullptr)                         return TTV_B200_ERR_PIC_NULL;
  if (na[q - 1] != nb[0])                     return TTV_B200_ERR_EXTENT_MISMATCH;
  if (!is_valid_shape(na, p))                 return TTV_B200_ERR_SHAPE_A;
  if (!is_valid_shape(nc, p - 1))             return TTV_B200_ERR_SHAPE_C;     // p == 1 always ends here
  if (!is_valid_layout(pia, p))               return TTV_B200_ERR_LAYOUT_A;
  if (!is_valid_layout(pic, p - 1))           return TTV_B200_ERR_LAYOUT_C;
  if (!is_valid_strides(pia, p, wa))          return TTV_B200_ERR_STRIDES_A;
  if (!is_valid_strides(pic, p - 1, wc))      return TTV_B200_ERR_STRIDES_C;
  if (p > (uint64_t)kMaxOrder)                return TTV_B200_ERR_OPTS;

  View v;
  v.ref_case = (uint32_t)classify_case(p, q, pia);
  uint64_t k = 0;
  while (pia[k] != q) ++k;                    // 0-based position of q in the layout
  v.k  = (uint32_t)(k + 1);
  v.nq = na[q - 1];
  for (uint64_t r = p; r-- > 0;)
    if (na[pia[r] - 1] > 1) { v.slow_extent = na[pia[r] - 1]; break; }
  for (uint64_t r = 0; r < k; ++r)     v.inner *= na[pia[r] - 1];
  for (uint64_t r = k + 1; r < p; ++r) v.outer *= na[pia[r] - 1];

  const bool honor = (flags & TTV_B200_FLAG_HONOR_STRIDES) != 0;
  if (v.ref_case == 8 || honor) {
    // C's layout must be A's layout without q (reference detail/tensor_times_vector.h:147-168).  In cases 1-7 the
    // reference never looks at pic, wa, wc: it runs one GEMV on the packed tensor and writes C packed in the derived
    // layout (detail/matrix_times_vector.h:314-336) -- and so does this library, unless TTV_B200_FLAG_HONOR_STRIDES
    // asks for wa / wc / pic to be taken at their word in every case (then C may even have a layout of its own).
    bool derived = true;                                   // pic is pia without q
    for (uint64_t i = 0; i < k; ++i)
      if (pic[i] != pia[i] - (pia[i] > q ? 1 : 0)) { if (!honor) return TTV_B200_ERR_LAYOUT_BEGIN; derived = false; }
    for (uint64_t i = k; i + 1 < p; ++i)
      if (pic[i] != pia[i + 1] - (pia[i + 1] > q ? 1 : 0)) { if (!honor) return TTV_B200_ERR_LAYOUT_END; derived = false; }

    // The loop nest of the reference walks A and C with wa / wc (tensor_times_vector.h:189-216) and asserts
    // wa[q-1] == inner in the subtensor variants (:956).  Packed strides are the documented input (README.md:41) and
    // take the fast kernels; any other valid strides are honoured the way the slice variants do, by the general-stride
    // kernel (strided_kernel.cuh).  Extent-1 modes may carry any stride.
    bool packed = derived;
    uint64_t expect = 1;
    for (uint64_t r = 0; r < p; ++r) {
      const uint64_t m = pia[r] - 1;
      if (na[m] > 1 && wa[m] != expect) packed = false;
      expect *= na[m];
    }
    expect = 1;
    for (uint64_t r = 0; r + 1 < p; ++r) {
      const uint64_t mc = pic[r] - 1;                 // mode of C
      const uint64_t ma = mc + (mc + 1 >= q ? 1 : 0); // the same mode in A
      if (na[ma] > 1 && wc[mc] != expect) packed = false;
      expect *= na[ma];
    }
    if (!packed) {
      v.strided = true;
      v.wq = wa[q - 1];
      v.span_a = 1; v.span_c = 1;
      for (uint64_t m = 0; m < p; ++m) v.span_a += (na[m] - 1) * wa[m];
      for (uint64_t r = 0; r + 1 < p; ++r) {
        const uint64_t mc = pic[r] - 1, ma = mc + (mc + 1 >= q ? 1 : 0);
        if (nc[mc] != na[ma]) return TTV_B200_ERR_NOT_PACKED;          // C's shape must be A's without mode q
        v.span_c += (nc[mc] - 1) * wc[mc];
        if (na[ma] == 1) continue;
        if (v.nfree > 0 && v.fwa[v.nfree - 1] * v.fn[v.nfree - 1] == wa[ma] && v.fwc[v.nfree - 1] * v.fn[v.nfree - 1] == wc[mc]) {
          v.fn[v.nfree - 1] *= na[ma];                                   // packed against its neighbour in A and in C
          continue;
        }
        if (v.nfree == (uint32_t)kMaxFreeDims) return TTV_B200_ERR_NOT_PACKED;
        v.fn[v.nfree] = na[ma]; v.fwa[v.nfree] = wa[ma]; v.fwc[v.nfree] = wc[mc];
        ++v.nfree;
      }
    }
  }
  *view = v;
  return TTV_B200_OK;
}

// ---- kernel chooser -----------------------------------------------------------------------------------------
static uint64_t ceil_div(uint64_t a, uint64_t b) { return (a + b - 1) / b; }
static uint64_t pow2_floor(uint64_t x) { uint64_t r = 1; while (r * 2 <= x) r *= 2; return r; }
static uint64_t pow2_ceil(uint64_t x) { uint64_t r = 1; while (r < x) r *= 2; return r; }

static uint64_t vmax_of(uint64_t s) { return s >= 16 ? 1 : 16 / s; }

// shared-memory bank conflict degree of the STREAM kernel's reads: lane u owns output (u / inner, u % inner) and reads
// word ((u / inner) * M + u % inner) * w of the stage (w = 32-bit words per element; wide accesses go out per
// 32/w lanes)
static unsigned stream_conflict_degree(uint64_t M, uint64_t inner, uint64_t s)
{
  const uint64_t w = std::max<uint64_t>(1, s / 4), lanes = 32 / std::min<uint64_t>(w, 4);
  unsigned count[32] = {0}, worst = 0;
  if (inner == 1 && M % 2 == 0) M += 1;          // even fibers are walked skewed (stream_fibers_skewed): odd lane stride
  for (uint64_t u = 0; u < lanes; ++u) {
    const uint64_t bank = (((u / inner) * M + (u % inner)) * w) % 32;
    worst = std::max(worst, ++count[bank]);
  }
  return worst;
}

// Experiment switches (TTV_B200_*) are read on every call so that a running process can flip them -- but one getenv per
// switch is ~30 scans of the environment per plan (2-3 us, as much as the launch itself).  choose_launch scans the
// environment ONCE, keeps the TTV_B200_ entries (normally none) and env_int looks names up in that short list.
extern "C" char** environ;

namespace {
struct EnvScan {
  static constexpr int kMax = 48;
  const char* entry[kMax];
  int n = 0;
  EnvScan()
  {
    for (char** e = environ; e && *e; ++e)
      if (std::strncmp(*e, "TTV_B200_", 9) == 0 && n < kMax) entry[n++] = *e;
  }
  const char* find(const char* name) const
  {
    const size_t len = std::strlen(name);
    for (int i = 0; i < n; ++i)
      if (std::strncmp(entry[i], name, len) == 0 && entry[i][len] == '=') return entry[i] + len + 1;
    return nullptr;
  }
};
thread_local const EnvScan* t_env = nullptr;
struct EnvScope {
  EnvScan scan;
  const EnvScan* prev;
  EnvScope() : prev(t_env) { t_env = &scan; }
  ~EnvScope() { t_env = prev; }
};
} // namespace

static int env_int(const char* name, int fallback)
{
  const char* s = t_env ? t_env->find(name) : std::getenv(name);
  return (s && *s) ? std::atoi(s) : fallback;
}

// COLX launch shape.  TY = V lanes along n_q (one per phase), TX lanes along inner; a tile owns W output columns and
// loads W/V + 1 vectors per row; W is chosen so that the tiles of a row are of (nearly) equal width.
static int choose_colx(uint64_t s, bool is_complex, const View& v, const ttv_b200_opts* opts, uint64_t sms, Launch* out)
{
  Launch l;
  const uint64_t V = vmax_of(s), NT = 256;
  uint64_t TY = (uint64_t)env_int("TTV_B200_COLX_TY", (int)V);
  if (TY < V || TY > 32 || TY % V) return TTV_B200_ERR_OPTS;
  const uint64_t txmax = NT / TY;
  int want = opts ? opts->ksplit : 0;
  if (want < 0) return TTV_B200_ERR_OPTS;
  if (want == 0) want = env_int("TTV_B200_KSPLIT", 0);

  l.kernel = TTV_B200_KERNEL_COLX;
  l.threads = (uint32_t)NT;
  l.vec = (int)V; l.ty = (uint32_t)TY; l.to = 1; l.udir = 0;
  l.stream = (uint32_t)env_int("TTV_B200_STREAM", 1);

  // Short contractions (fewer than 48 rows per phase lane): the per-tile prologue / shared-memory epilogue of the CTA
  // form dominates (measured 23^7 q=4: 4.0 TB/s against 6.4 TB/s; 73^5 fp64 q=3,5, 37 rows per lane: 6.0 against 6.7), so
  // the warp-autonomous form runs them; long contractions are faster in the CTA form (1625^3 q=2: 7.3 against 6.8 TB/s;
  // 215^4 fp64, 108 rows per lane: 6.9), which keeps 3 CTAs per SM.
  const int warp_mode = env_int("TTV_B200_COLX_WARP", -1);      // -1 auto, 0 CTA form, 1 COLW (phase classes), 2 COLR (realigned)
  if (warp_mode == 1 || warp_mode == 2 || (warp_mode == -1 && ceil_div(v.nq, TY) < (uint64_t)env_int("TTV_B200_COLX_WARP_ROWS", 48))) {
    // a warp owns 31*V columns per unit and all rows of its n_q partition
    // Measured (23^7 q=4,7 fp32; 21^7 q=7 fp64): the realigned form (80 registers, 3 CTAs per SM, whole-sector stores)
    // is no faster than the phase-class form (115 registers, 2 CTAs): 6.48-6.53 against 6.46-6.55 TB/s in fp32, 6.33
    // against 6.57 in fp64 -- neither residency nor the partial-sector stores bound these shapes -- so COLW stays the
    // default and COLR is kept selectable (profiles/r01_colr_probe.txt).  EXCEPT complex<float> (round 2, 25^6 in three
    // layouts, profiles/r02_variants_weak_shapes.txt): COLW keeps V x V complex accumulators per unit there and runs
    // 5 863-6 388 GB/s where COLR runs 6 394-6 661 ([15625, 25, 625]: 5 863 -> 6 394), so complex<float> takes COLR.
    const bool colr = warp_mode == 2 || (warp_mode == -1 && is_complex && s == 8 && env_int("TTV_B200_COLR_C64", 1) != 0);
    l.warp = colr ? 2 : 1; l.tx = 32; l.ty = 1;
    uint64_t Vw = V, loads = 8;
    uint64_t ku = V;
    if (l.warp == 1) {
      // COLW, measured (23^7 q=4,7 fp32; 21^7 q=7 fp64): two rows in flight for many units beat deeper batches (6.5
      // against 6.3 TB/s), and 4-byte elements do best as 8-byte vectors (2 phases: 4 accumulators per unit, not 16).
      if (s == 4 && env_int("TTV_B200_COLW_V", 2) == 2) { Vw = 2; loads = 16; l.vec = 2; }
      ku = Vw == 2 ? 2 : 4;
    } else {
      // COLR: full 16-byte vectors, one phase period of rows per batch
      ku = V;
    }
    const int ku_env = env_int("TTV_B200_KU", 0);
    if ((ku_env == 2 || ku_env == 4 || ku_env == 8) && (uint64_t)ku_env % Vw == 0) ku = (uint64_t)ku_env;
    uint64_t nu = loads / ku;
    l.ku = (int)ku; l.nu = (int)nu;
    l.wcols = 31 * Vw;
    l.itiles = ceil_div(v.inner, nu * l.wcols);
    l.otiles = v.outer;
    const uint64_t tiles1 = l.itiles * l.otiles;
    uint64_t ksplit = 1;
    if (want > 0) ksplit = (uint64_t)want;
    else if (tiles1 < sms * 8) ksplit = std::min(ceil_div(sms * 64, tiles1), std::max<uint64_t>(1, v.nq / (ku * 4)));
    ksplit = std::max<uint64_t>(1, std::min(ksplit, v.nq));
    const uint64_t kchunk = ceil_div(v.nq, ksplit);
    ksplit = ceil_div(v.nq, kchunk);
    l.ksplit = (uint32_t)ksplit; l.kchunk = kchunk;
    l.tiles = tiles1 * ksplit;
    l.ctas = std::min<uint64_t>(ceil_div(l.tiles, NT / 32), sms * 32);
    l.kb = 0;
    l.smem_bytes = l.warp == 2 ? (NT / 32) * nu * l.wcols * s : 0;      // COLR: one output strip per warp
    l.workspace_bytes = ksplit > 1 ? ksplit * v.outer * v.inner * s : 0;
    *out = l;
    return TTV_B200_OK;
  }

  // batch shape: 8 vector loads in flight per thread; short contractions take the depth that wastes least
  const uint64_t per = ceil_div(v.nq, TY);
  auto waste = [per](uint64_t d) { return (double)(ceil_div(per, d) * d - per) / (double)(ceil_div(per, d) * d); };
  uint64_t ku = 8;
  if (per >= 16) { while (ku > 2 && waste(ku) > 0.10) ku /= 2; }
  else { ku = 2; for (uint64_t d = 4; d <= 8; d *= 2) if (waste(d) < waste(ku) - 1e-9) ku = d; }
  const int ku_env = env_int("TTV_B200_KU", 0);
  if (ku_env == 2 || ku_env == 4 || ku_env == 8) ku = (uint64_t)ku_env;
  uint64_t nu = 8 / ku;
  // units widen a tile: never starve the SMs of tiles for their sake
  while (nu > 1 && ceil_div(v.inner, (txmax * nu - 1) * V) * v.outer < sms * 8) { nu /= 2; ku *= 2; }
  l.ku = (int)ku; l.nu = (int)nu;

  const uint64_t wmax = (txmax * nu - 1) * V;
  l.itiles = ceil_div(v.inner, wmax);
  l.wcols = ceil_div(ceil_div(v.inner, l.itiles), V) * V;              // <= wmax, multiple of V
  l.tx = (uint32_t)ceil_div(l.wcols / V + 1, nu);                      // tx*nu vectors cover W columns at any phase
  l.otiles = v.outer;
  l.a_ustride = (uint64_t)l.tx * V;
  l.c_ustride = (uint64_t)l.tx * V;

  const uint64_t tiles1 = l.itiles * l.otiles, kstep = TY;
  uint64_t ksplit = 1;
  if (want > 0) ksplit = (uint64_t)want;
  else if (tiles1 < sms) ksplit = std::min(ceil_div(sms * 8, tiles1), std::max<uint64_t>(1, v.nq / (kstep * 16)));
  ksplit = std::max<uint64_t>(1, std::min(ksplit, ceil_div(v.nq, kstep)));
  const uint64_t kchunk = ceil_div(ceil_div(v.nq, ksplit), kstep) * kstep;
  ksplit = ceil_div(v.nq, kchunk);
  l.ksplit = (uint32_t)ksplit; l.kchunk = kchunk;
  l.tiles = tiles1 * ksplit;
  l.ctas = std::min<uint64_t>(l.tiles, sms * 64);

  uint64_t kb = std::min<uint64_t>(kchunk, 16384 / s);
  kb = std::max<uint64_t>(kstep, kb / kstep * kstep);
  l.kb = (uint32_t)kb;
  l.smem_bytes = kb * s + TY * l.wcols * s;
  l.workspace_bytes = ksplit > 1 ? ksplit * v.outer * v.inner * s : 0;
  *out = l;
  return TTV_B200_OK;
}

int choose_launch(int dtype, const View& v, const ttv_b200_opts* opts, uint64_t align_a, uint64_t align_b,
                  uint64_t align_c, int sm_count, Launch* out)
{
  const uint64_t s = (uint64_t)dtype_size(dtype);
  if (s == 0) return TTV_B200_ERR_DTYPE;
  if (sm_count <= 0) sm_count = 148;
  const uint64_t sms = (uint64_t)sm_count;
  const EnvScope env_scope;                       // one scan of the environment for all the switches below
  const uint32_t flags = opts ? opts->flags : 0u;
  int forced = opts ? opts->kernel : 0;
  if (forced < 0 || forced >= TTV_B200_KERNEL_COUNT || forced == TTV_B200_KERNEL_STRIDED) return TTV_B200_ERR_OPTS;
  Launch l;
  // COLT: A through shared memory by TMA tensor tiles (colt_kernel.cuh).  Needs rows that are whole 16-byte vectors (the
  // tensor map's strides), a 16-byte aligned A and C, b resident in shared memory, and tensor-map extents below 2^32.
  // Only taken when forced (opts.kernel / TTV_B200_USE_COLT=1): measured against COL it does not win.
  {
    const uint64_t row_words = v.inner * s / 4;
    const bool eligible = v.inner > 1 && (v.inner * s) % 16 == 0 && (align_a % 16) == 0 && (align_c % 16) == 0 && row_words >= 16 &&
                          row_words < (1ull << 32) && v.nq < (1ull << 32) && v.outer < (1ull << 32) && v.nq * s <= 64 * 1024 &&
                          v.nq * v.inner * s < (1ull << 40) && !(flags & TTV_B200_FLAG_NO_VEC);
    if (forced == TTV_B200_KERNEL_COLT && !eligible) return TTV_B200_ERR_OPTS;
    const bool pick = forced == TTV_B200_KERNEL_COLT || (forced == 0 && eligible && env_int("TTV_B200_USE_COLT", 0) == 1);
    if (pick) {
      l.kernel = TTV_B200_KERNEL_COLT;
      l.threads = 288;                                              // 256 consumers + the producer warp
      l.vec = (int)(16 / s); l.nu = 1; l.stream = 1; l.udir = 0;
      uint64_t wt = std::min<uint64_t>(256, pow2_ceil(row_words));   // words of a row per box (power of two: tx divides 256)
      wt = std::max<uint64_t>(4, std::min<uint64_t>(wt, (uint64_t)env_int("TTV_B200_COLT_WT", 256)));
      l.wt = (uint32_t)wt;
      l.tx = (uint32_t)(wt / 4); l.ty = 256 / l.tx; l.to = 1;
      const uint64_t stage_target = (uint64_t)env_int("TTV_B200_COLT_STAGE_KB", 32) * 1024;
      uint64_t kt = std::max<uint64_t>(l.ty, std::min<uint64_t>(256, stage_target / (wt * 4)) / l.ty * l.ty);
      kt = std::min<uint64_t>(kt, ceil_div(v.nq, l.ty) * l.ty);      // never taller than the contraction
      l.kt = (uint32_t)kt;
      l.ku = (int)(kt / l.ty);
      int want = opts ? opts->ksplit : 0;
      if (want == 0) want = env_int("TTV_B200_KSPLIT", 0);
      const uint64_t boxes_all = ceil_div(v.nq, kt);
      uint64_t ksplit = want > 0 ? (uint64_t)want : 1;
      ksplit = std::max<uint64_t>(1, std::min(ksplit, boxes_all));
      const uint64_t boxes_per = ceil_div(boxes_all, ksplit);
      ksplit = ceil_div(boxes_all, boxes_per);
      l.ksplit = (uint32_t)ksplit;
      l.kchunk = boxes_per * kt;
      l.kboxes = (uint32_t)boxes_per;
      l.itiles = ceil_div(row_words, wt);
      l.otiles = v.outer;
      l.tiles = l.itiles * l.otiles * ksplit;
      l.stages = (uint32_t)std::max(2, std::min(8, env_int("TTV_B200_COLT_STAGES", 4)));
      const uint64_t b_bytes = (boxes_per * ksplit * kt * s + 15) / 16 * 16;  // all of b, zero-padded to whole boxes
      l.kb = (uint32_t)(boxes_per * ksplit * kt);
      l.smem_bytes = (uint64_t)l.stages * wt * kt * 4 + b_bytes + 256 * 16 + 2 * (uint64_t)l.stages * 8 + 128;
      if (l.smem_bytes > 227 * 1024) return TTV_B200_ERR_OPTS;
      const uint64_t per_sm = std::max<uint64_t>(1, std::min<uint64_t>((227 * 1024) / (l.smem_bytes + 1024), (uint64_t)env_int("TTV_B200_COLT_CTAS", 2)));
      l.ctas = std::min<uint64_t>(l.tiles, sms * per_sm);
      l.workspace_bytes = ksplit > 1 ? ksplit * v.outer * v.inner * s : 0;
      *out = l;
      return TTV_B200_OK;
    }
  }
  // DOTP: fibers of two elements of 4 or 8 bytes (dotp_kernel.cuh): whole fibers per 16-byte vector, b in registers, one
  // 8-byte store per vector, nothing shared.  STREAM took the 4-byte shapes before ([1610612736, 2, 1]: 6 311-6 539 GB/s,
  // now 7 002-7 060), DOTF the 8-byte ones ([536870912, 2, 1] fp64 6 738 -> 6 970) -- tools/probe/tiny_inner.py.
  {
    const bool eligible = v.inner == 1 && v.nq == 2 && s <= 8 && (align_a % 16) == 0 && (align_c % 8) == 0 && (!opts || opts->ksplit <= 1) &&
                          !(flags & TTV_B200_FLAG_NO_VEC);
    if (forced == TTV_B200_KERNEL_DOTP && !eligible) return TTV_B200_ERR_OPTS;
    const int mode = env_int("TTV_B200_USE_DOTP", -1);                                 // -1 auto, 0 never, 1 whenever eligible
    const bool pick = forced == TTV_B200_KERNEL_DOTP ? true
                    : forced != 0 ? false
                    : mode == 1 ? eligible
                    : mode == 0 ? false
                    : eligible;
    if (pick) {
      l.kernel = TTV_B200_KERNEL_DOTP;
      l.threads = 256;
      l.vec = (int)(16 / s); l.tx = 1; l.ty = 1; l.to = 256; l.nu = 1; l.stream = 1; l.udir = 0; l.ksplit = 1;
      l.ku = env_int("TTV_B200_DOTP_KU", 8) == 4 ? 4 : 8;
      const uint64_t nvec = v.outer * 2 * s / 16;
      l.tiles = std::max<uint64_t>(1, ceil_div(nvec, 256ull * (uint64_t)l.ku));
      // one tile per CTA: measured 6 432 (64 CTAs per SM striding over the tiles) -> 7 000 GB/s on [1610612736, 2, 1]
      l.ctas = std::min<uint64_t>(l.tiles, std::min<uint64_t>(0x7fffffffull, sms * (uint64_t)std::max(1, env_int("TTV_B200_DOTP_CTAS", 1 << 20))));
      l.kchunk = v.nq; l.kb = 2;
      l.smem_bytes = 0;
      l.workspace_bytes = 0;
      *out = l;
      return TTV_B200_OK;
    }
  }
  // COLF, tiny slabs: n_q = 2 .. 32 rows of two 4-byte elements (a slab is 1 .. 16 vectors): consecutive lanes on consecutive
  // vectors, transposing butterfly over the lanes of a slab (ttv_colf_tiny_kernel).  STREAM / COL / the lane-per-slab form of
  // COLF took these before ([16777216, 16, 2]: COL 1 814, COLF 4 729 GB/s).
  {
    const bool eligible = s == 4 && v.inner == 2 && v.nq >= 2 && v.nq <= 32 && (v.nq & (v.nq - 1)) == 0 && (align_a % 16) == 0 && (align_c % 8) == 0 &&
                          (!opts || opts->ksplit <= 1) && !(flags & TTV_B200_FLAG_NO_VEC);
    const int mode = env_int("TTV_B200_COLF_TINY", -1);                                // -1 auto, 0 never, 1 whenever eligible
    const bool pick = (forced == TTV_B200_KERNEL_COLF || forced == 0) && eligible && mode != 0 && env_int("TTV_B200_USE_COLF", -1) != 0;
    if (pick) {
      l.kernel = TTV_B200_KERNEL_COLF;
      l.tiny = 1;
      l.threads = 256;
      const uint64_t G = v.nq / 2;
      l.vec = 4; l.tx = 1; l.ty = (uint32_t)G; l.to = 2; l.nu = (int)(32 / G); l.ku = 8; l.stream = 1; l.udir = 0; l.ksplit = 1;
      l.kchunk = v.nq; l.kb = 0;
      l.itiles = 1; l.otiles = ceil_div(v.outer, 8 * (32 / G));
      l.tiles = l.otiles;
      // a CTA per eight items, no striding (a lane's set-up is two elements of b): 8 CTAs per SM striding over the items measured
      // 6 297-6 630 GB/s, 32 per SM 6 608-6 873, a CTA per eight items 7 002-7 081 (profiles/r02_colf_dotp_ab.txt, section 16)
      l.ctas = std::min<uint64_t>(ceil_div(l.tiles, 8), std::min<uint64_t>(0x7fffffffull, sms * (uint64_t)std::max(1, env_int("TTV_B200_COLF_TINY_CTAS", 1 << 20))));
      l.smem_bytes = 0;
      l.workspace_bytes = 0;
      *out = l;
      return TTV_B200_OK;
    }
  }

  // STREAM: small slabs staged through shared memory by TMA bulk copies (stream_kernel.cuh).  Eligible when a slab and
  // b are small, A is 16-byte aligned and n_q is not split.
  {
    const uint64_t slab_bytes = v.nq * v.inner * s;
    // Small slabs share a stage of up to 36 KB (2 CTAs x 3 stages per SM).  A slab of up to 75 KB gets a stage of its own
    // (1 CTA of up to 1024 threads x 3 stages per SM, ~100-150 KB in flight): 23 x 529 floats (48.7 KB) run at 6.7-6.8
    // TB/s this way, 4.9 with the warp form of COLX, whose units of 62 columns tile a row of 529 badly; 21 x 441 doubles
    // (74 KB) 6.5 against 5.9.
    // Stage size of the shared (small-slab) form, measured per shape family in round 2 (profiles/r02_stream_stage_sweep.txt;
    // the landscape is not monotonic): 36 KB x 3 stages x 2 CTAs per SM for 4-byte elements and for n_q = 2, 3; FIBERS of
    // 8-byte elements do better with one CTA per SM and up to 56 KB per stage (21 doubles 6 516 -> 6 939, 73 doubles
    // 6 387 -> 7 045, 25 complex<float> 6 146 -> 6 767 GB/s), slabs of 8-byte elements of up to 4 KB with 24 KB stages
    // (21 x 21 doubles 6 408 -> 6 874).
    int stage_kb = 36;
    if (s == 8 && v.nq >= 8) stage_kb = v.inner == 1 ? 56 : (slab_bytes <= 4096 ? 24 : 36);
    const uint64_t small_payload = (uint64_t)env_int("TTV_B200_STAGE_KB", stage_kb) * 1024 - 128;
    const uint64_t big_slab_max = (uint64_t)env_int("TTV_B200_STREAM_SLAB_KB", 75) * 1024;
    const bool big_slab = slab_bytes > small_payload;
    const uint64_t big_stage = (slab_bytes + 32 + 127) / 128 * 128, b_bytes16 = (v.nq * s + 15) / 16 * 16;
    const uint64_t big_smem = 3 * big_stage + b_bytes16 + 3 * 8;                        // at least three stages of one slab
    const bool eligible = slab_bytes <= std::max<uint64_t>(8192, big_slab_max) && (!big_slab || big_smem <= 227 * 1024) &&
                          v.nq * s <= 8192 && (align_a % 16) == 0 && (!opts || opts->ksplit <= 1) && v.outer * v.nq * v.inner * s >= 16;
    const uint64_t max_payload = big_slab ? slab_bytes : small_payload;
    // outputs per stage against the 256 threads of the CTA (measured: fibers of 73 doubles, 62 per stage: 6.4 TB/s
    // against 5.5 with the peeled DOT kernel)
    const bool fills_cta = (max_payload / std::max<uint64_t>(1, slab_bytes)) * v.inner >= (uint64_t)env_int("TTV_B200_STREAM_MIN_OUT", 48);
    const bool misaligned = (v.inner == 1 ? v.nq : v.inner) % vmax_of(s) != 0;      // the other kernels would fall back to narrow loads
    const int mode = env_int("TTV_B200_USE_STREAM", -1);                               // -1 auto, 0 never, 1 whenever eligible
    const bool pick = forced == TTV_B200_KERNEL_STREAM ? eligible
                    : forced != 0 ? false
                    : mode == 1 ? eligible
                    : mode == 0 ? false
                    // (a big slab needs enough outputs for the thread-per-output mapping: 73 x 73 doubles, 73 outputs,
                    // ran 5.7 TB/s against 6.6 with the column kernel)
                    : (eligible && misaligned && fills_cta && (!big_slab || v.inner >= 256) && v.outer >= sms * 64 && !(flags & TTV_B200_FLAG_NO_VEC) &&
                       stream_conflict_degree(v.nq * v.inner, v.inner, s) <= 2);
    if (forced == TTV_B200_KERNEL_STREAM && !eligible) return TTV_B200_ERR_OPTS;
    if (pick) {
      // stage size: up to 36 KB (3 stages x 2 CTAs per SM = 216 KB of shared memory, all of it in flight), trimmed so
      // that the outputs of a chunk fill whole rounds of the CTA's 256 threads (measured: 2 CTAs x 24-36 KB is best)
      const uint64_t payload = max_payload;
      l.kernel = TTV_B200_KERNEL_STREAM;
      l.threads = (uint32_t)env_int("TTV_B200_STREAM_THREADS", 256);
      // a big slab alone in its stage: as many threads as it has outputs (one round), up to 1024, fibers excepted
      if (big_slab && v.inner > 256 && env_int("TTV_B200_STREAM_THREADS", 0) == 0)
        l.threads = (uint32_t)std::min<uint64_t>(1024, ceil_div(v.inner, 128) * 128);
      l.vec = 1; l.tx = 1; l.ty = 1; l.to = 1; l.nu = 4; l.ku = 1; l.ksplit = 1; l.stream = 1;
      uint64_t S = std::max<uint64_t>(1, std::min<uint64_t>(payload / slab_bytes, v.outer));
      const uint64_t rounds = S * v.inner / l.threads;
      if (rounds >= 1) S = std::max<uint64_t>(1, rounds * l.threads / v.inner);
      l.slabs_per_chunk = S;
      l.chunks = ceil_div(v.outer, l.slabs_per_chunk);
      l.stage_bytes = (uint32_t)((l.slabs_per_chunk * slab_bytes + 32 + 127) / 128 * 128);
      l.tiles = l.chunks;
      // a big slab is alone in its stage and the CTA alone on its SM; more than three stages (TTV_B200_STREAM_STAGES, up
      // to 5 when they fit) measured no better: 23 x 529 floats 6.80 TB/s with 3 stages, 6.66 with 4
      l.stages = 3;
      l.stages = (uint32_t)std::max<uint64_t>(3, std::min<uint64_t>(std::min<uint64_t>(5, (uint64_t)env_int("TTV_B200_STREAM_STAGES", 3)), (227 * 1024 - b_bytes16 - 64) / l.stage_bytes));
      l.smem_bytes = (uint64_t)l.stages * l.stage_bytes + b_bytes16 + (uint64_t)l.stages * 8;
      if (l.smem_bytes > 227 * 1024) return TTV_B200_ERR_OPTS;                          // (excluded by `eligible`)
      const uint64_t per_sm = l.smem_bytes + 1024 <= (227 * 1024) / 2 ? 2 : 1;        // CTAs of this size an SM can hold
      l.ctas = std::min<uint64_t>(l.chunks, sms * std::min<uint64_t>(per_sm, (uint64_t)env_int("TTV_B200_STREAM_CTAS", 2)));
      l.kchunk = v.nq; l.kb = (uint32_t)v.nq;
      l.workspace_bytes = 0;
      *out = l;
      return TTV_B200_OK;
    }
    if (forced == TTV_B200_KERNEL_STREAM) forced = 0;
  }

  // COLF: rows that are not whole 16-byte vectors, streamed flat as super-rows of R = V / gcd(inner, V) rows =
  // L = inner / gcd whole vectors, a warp per slab or slab partition (colf_kernel.cuh).  Slabs must start on a vector
  // boundary (n_q a multiple of R, or one slab whose last rows the kernel takes with plain loads).
  {
    const uint64_t Vf = vmax_of(s);
    uint64_t g = Vf, t = v.inner % Vf;
    while (t) { const uint64_t r = g % t; g = t; t = r; }             // gcd(inner, V)
    const uint64_t R = Vf / g, L = v.inner / g;
    const bool eligible = Vf > 1 && v.inner > 1 && R > 1 && L <= 32 && (v.nq % R == 0 || v.outer == 1) && (align_a % 16) == 0 &&
                          !(flags & TTV_B200_FLAG_NO_VEC);
    // measured (tools/probe/tiny_inner.py, profiles/r02_colf_dotp_ab.txt): 4-byte elements win at every L <= 32 from slabs of 128
    // bytes on (64-byte slabs lose: 3 347 against 4 235 GB/s); 8-byte elements, whose column kernel already loads 8 bytes
    // per lane, win with rows of 3 / 5 / 7 and lose with rows of 21 (5 957 against 6 360)
    const bool pays = v.nq * v.inner * s >= (uint64_t)env_int("TTV_B200_COLF_MIN_SLAB_B", 128) && (s == 4 || L <= 8);
    if (forced == TTV_B200_KERNEL_COLF && !eligible) return TTV_B200_ERR_OPTS;
    const int mode = env_int("TTV_B200_USE_COLF", -1);                                 // -1 auto, 0 never, 1 whenever eligible
    const bool pick = forced == TTV_B200_KERNEL_COLF ? true
                    : forced != 0 ? false
                    : mode == 1 ? eligible
                    : mode == 0 ? false
                    : (eligible && pays);
    if (pick) {
      constexpr uint64_t KUf = 8;
      l.kernel = TTV_B200_KERNEL_COLF;
      l.threads = 256;
      const uint64_t nsr = v.nq / R, ty_max = 32 / L;                // super-rows of a slab; lanes per phase when a warp has one slab
      int want = opts ? opts->ksplit : 0;
      if (want < 0) return TTV_B200_ERR_OPTS;
      if (want == 0) want = env_int("TTV_B200_KSPLIT", 0);
      // short slabs: fewer lanes per slab, so that a lane still has a batch of loads, and several slabs side by side in a warp
      uint64_t ty = want > 1 ? ty_max : std::min(ty_max, std::max<uint64_t>(1, nsr / KUf));
      const uint64_t sw = ty < ty_max ? 32 / (ty * L) : 1;
      if (sw == 1) ty = ty_max;
      l.vec = (int)Vf; l.tx = (uint32_t)L; l.ty = (uint32_t)ty; l.to = (uint32_t)R; l.nu = (int)sw; l.ku = (int)KUf; l.udir = 0;
      l.stream = ty * L * 16 >= 128 ? 1 : 0;                         // lane groups narrower than a line reuse it from L1
      l.pair = (v.inner == 2 && s == 4 && (align_b % 8) == 0 && (align_c % 8) == 0 && env_int("TTV_B200_COLF_PAIR", 1) != 0) ? 1 : 0;
      const uint64_t batch = ty * KUf;                               // super-rows of one batch of a lane group
      // persistent warps striding over the items (slab group, partition): twice the CTAs an SM holds
      const uint64_t resident = s == 8 ? 2 : 3;
      const uint64_t max_ctas = sms * (uint64_t)std::max(1, env_int("TTV_B200_COLF_CTAS", (int)(2 * resident)));
      const uint64_t vslots = max_ctas * 8;
      const uint64_t ogroups = ceil_div(v.outer, sw);
      // n_q split: up to four items per warp (more partitions measured slower: [64, 2^19, 7] 216 partitions 6 026, 820
      // 5 784 GB/s -- the reduce pass is a second launch), at least four batches per partition; among the candidates
      // the one whose item count fills its rounds best (a last round that is 4 % full costs a whole round)
      uint64_t ksplit = 1;
      auto parts = [&](uint64_t ks) { const uint64_t chunk = ceil_div(std::max<uint64_t>(1, ceil_div(nsr, ks)), batch) * batch; return std::max<uint64_t>(1, ceil_div(nsr, chunk)); };
      if (want > 0) ksplit = parts(std::min<uint64_t>((uint64_t)want, std::max<uint64_t>(1, nsr)));
      else if (sw == 1) {
        // (never more than 24 partitions per SM: every partition of an output is one more term of the reduce pass, which
        // for one or a few slabs is a handful of CTAs -- [1, 2^26, 2] 13 108 partitions 4 557, 3 543 partitions 6 458 GB/s)
        const uint64_t hi = std::max<uint64_t>(1, std::min(std::min(ceil_div(vslots * (uint64_t)env_int("TTV_B200_COLF_ITEMS_PER_WARP", 4), ogroups), sms * 24),
                                                           std::max<uint64_t>(1, nsr / (batch * 4))));
        double best = -1.0;
        const uint64_t lo = std::max<uint64_t>(1, hi / 2), step = std::max<uint64_t>(1, (hi - lo) / 64);   // at most ~64 candidates: this runs per call
        for (uint64_t ks = hi; ks >= lo; ks -= std::min(step, ks - lo ? ks - lo : step)) {
          const uint64_t k2 = parts(ks);
          const double x = (double)(ogroups * k2) / (double)vslots, eff = x / (double)ceil_div(ogroups * k2, vslots);
          if (eff > best + 1e-9) { best = eff; ksplit = k2; }
          if (ks == lo) break;
        }
      }
      const uint64_t srchunk = ceil_div(std::max<uint64_t>(1, ceil_div(nsr, ksplit)), batch) * batch;
      l.ksplit = (uint32_t)ksplit; l.kchunk = srchunk * R;
      l.short1 = (ksplit == 1 && nsr <= batch && env_int("TTV_B200_COLF_SHORT", 1) != 0) ? 1 : 0;
      // the short-slab form of two-element rows sets a lane up with eight pairs of b only: more, shorter-lived CTAs measured
      // faster (6 per SM 6 691-6 768, 16 per SM 6 869-6 915, 32 per SM 7 026-7 079 GB/s); the general short form loses with them
      uint64_t grid_cap = max_ctas;
      if (l.short1 && l.pair) grid_cap = sms * (uint64_t)std::max(1, env_int("TTV_B200_COLF_PAIR_CTAS", 32));
      l.itiles = 1; l.otiles = ogroups;
      l.tiles = ogroups * ksplit;
      l.ctas = std::min<uint64_t>(ceil_div(l.tiles, 8), grid_cap);
      l.kb = 0;
      l.smem_bytes = 8 * 32 * 16;                                    // static: a strip of 32 vectors per warp
      l.workspace_bytes = ksplit > 1 ? ksplit * v.outer * v.inner * s : 0;
      *out = l;
      return TTV_B200_OK;
    }
  }

  // STREAMK: a tiny inner extent that no 16-byte vector tiles (3, 5, 6, 7, 9 ... elements) under a long contraction
  // (streamk_kernel.cuh).  The column kernel runs such rows with 4- / 8-byte loads and its lanes along n_q: measured
  // [64, 2^20, 3] fp32 4.3, [262144, 256, 3] 2.9, [48, 2^19, 5] fp64 5.9 TB/s (tools/probe/tiny_inner.py).
  {
    const uint64_t Vk = vmax_of(s);
    const bool eligible = v.inner > 1 && v.inner <= 16 && v.inner * s <= 64 && v.nq >= 4096 && (align_a % 16) == 0 && (align_b % 16) == 0 &&
                          !(flags & TTV_B200_FLAG_NO_VEC);
    if (forced == TTV_B200_KERNEL_STREAMK && !eligible) return TTV_B200_ERR_OPTS;
    const int mode = env_int("TTV_B200_USE_STREAMK", -1);
    const bool misaligned = (v.inner % Vk) != 0;
    const bool pick = forced == TTV_B200_KERNEL_STREAMK ? true
                    : forced != 0 ? false
                    : mode == 1 ? eligible
                    : mode == 0 ? false
                    // odd rows of 4-byte elements only: there the column kernel is down to 4-byte loads (rows of 3 / 5 floats off the
                    // 16-byte grid: 4 274 / 4 336 -> 5 810 / 5 963 GB/s); even rows and 8-byte elements it loads 8 bytes at a time and
                    // is ahead or level (rows of 2 floats 6 191 against 5 600, 3 doubles 5 694 / 5 811) -- tools/probe/streamk_niche.py
                    : (eligible && misaligned && s == 4 && (v.inner & 1) && v.outer * v.nq * v.inner * s >= (64ull << 20));
    if (pick) {
      l.kernel = TTV_B200_KERNEL_STREAMK;
      l.threads = 256;
      l.vec = 1; l.tx = 1; l.ty = 256; l.to = 1; l.nu = 1; l.ku = 1; l.stream = 1; l.udir = 0;
      const uint64_t budget = (uint64_t)env_int("TTV_B200_STREAMK_STAGE_KB", 32) * 1024;
      uint64_t rows = std::max<uint64_t>(256, budget / ((v.inner + 1) * s) / 256 * 256);
      rows = std::min<uint64_t>(rows, ceil_div(v.nq, 256) * 256);
      l.slabs_per_chunk = rows;
      l.stage_bytes = (uint32_t)((rows * v.inner * s + 32 + 127) / 128 * 128);
      l.b_stage_bytes = (uint32_t)((rows * s + 32 + 127) / 128 * 128);
      l.stages = 3;
      int want = opts ? opts->ksplit : 0;
      if (want < 0) return TTV_B200_ERR_OPTS;
      if (want == 0) want = env_int("TTV_B200_KSPLIT", 0);
      uint64_t ksplit = want > 0 ? (uint64_t)want : std::max<uint64_t>(1, ceil_div(sms * 8, v.outer));
      ksplit = std::max<uint64_t>(1, std::min(ksplit, std::max<uint64_t>(1, v.nq / (rows * 4))));     // four stages per partition at least
      if (want > 0) ksplit = std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)want, ceil_div(v.nq, rows)));
      uint64_t kchunk = ceil_div(ceil_div(v.nq, ksplit), rows) * rows;
      ksplit = ceil_div(v.nq, kchunk);
      l.ksplit = (uint32_t)ksplit; l.kchunk = kchunk;
      l.itiles = 1; l.otiles = v.outer;
      l.tiles = v.outer * ksplit;
      l.ctas = std::min<uint64_t>(l.tiles, sms * 2);
      l.kb = (uint32_t)rows;
      l.smem_bytes = 3ull * (l.stage_bytes + l.b_stage_bytes) + ((256 * v.inner * s + 15) / 16 * 16) + 3 * 8 + 64;
      if (l.smem_bytes > 113 * 1024) return TTV_B200_ERR_OPTS;
      l.workspace_bytes = ksplit > 1 ? ksplit * v.outer * v.inner * s : 0;
      *out = l;
      return TTV_B200_OK;
    }
  }

  // DOTF: short contiguous fibers read as one flat stream (dotf_kernel.cuh).  Measured against the lane-group DOT kernel:
  // 40 floats 7.0 against 6.0 TB/s, 84 floats 6.95 / 6.4, 40 complex<double> 6.4 / 5.5, 4 floats 6.9 / 6.4; from 64
  // vectors per fiber on the lane groups are as good or better (256 floats: 6.9 / 7.2), so they keep those.
  {
    const uint64_t Vf = vmax_of(s);
    const uint64_t nvf = v.nq / Vf;
    const bool eligible = v.inner == 1 && v.nq % Vf == 0 && nvf >= 1 && nvf <= 256 && (align_a % 16) == 0 && (align_b % 16) == 0 &&
                          (!opts || opts->ksplit <= 1) && !(flags & TTV_B200_FLAG_NO_VEC);
    if (forced == TTV_B200_KERNEL_DOTF && !eligible) return TTV_B200_ERR_OPTS;
    const int mode = env_int("TTV_B200_USE_DOTF", -1);
    const bool pick = forced == TTV_B200_KERNEL_DOTF ? true
                    : forced != 0 ? false
                    : mode == 1 ? eligible
                    : mode == 0 ? false
                    // (16-byte elements: up to 128 per fiber -- 128 complex<double> run 6.95 TB/s here, 6.18 with lane groups)
                    : (eligible && nvf <= (uint64_t)env_int("TTV_B200_DOTF_NV", s >= 16 ? 128 : 48) && v.outer >= sms * 64);
    if (pick) {
      l.kernel = TTV_B200_KERNEL_DOTF;
      l.threads = 256;
      l.vec = (int)Vf; l.tx = 1; l.ty = 1; l.to = 1; l.nu = 1; l.ku = 8; l.ksplit = 1; l.stream = 0; l.udir = 1;
      l.slabs_per_chunk = 256 / nvf;                                  // whole fibers per chunk of 256 vectors
      l.chunks = ceil_div(v.outer, l.slabs_per_chunk);
      l.tiles = l.chunks;
      // CTAs per SM the grid is capped at (session 33, profiles/r02_dotf_grid_ab.txt): 64 instead of 24 gains 1-2 % for 4- and 8-byte
      // elements (84 floats 6 921 -> 7 070, 48 complex<float> 6 732 -> 6 867, 16 floats 6 836 -> 6 989), loses up to 3 % for 16-byte
      // ones (40 complex<double> 6 438 -> 6 242); a CTA per eight chunks loses 8-40 % everywhere
      l.ctas = std::min<uint64_t>(ceil_div(l.chunks, 8), std::min<uint64_t>(0x7fffffffull, sms * (uint64_t)std::max(1, env_int("TTV_B200_DOTF_CTAS", s >= 16 ? 24 : 64))));
      l.kchunk = v.nq; l.kb = (uint32_t)v.nq;
      l.smem_bytes = v.nq * s + 8 * 256 * s;
      l.workspace_bytes = 0;
      *out = l;
      return TTV_B200_OK;
    }
  }

  // COLX: wide rows that start off 16-byte boundaries (odd inner extent): phase lanes along n_q keep the loads at
  // 16 bytes (colx_kernel.cuh).  Needs a 16-byte aligned A and an element type narrower than 16 bytes.
  {
    const uint64_t Vx = vmax_of(s);
    const bool eligible = Vx > 1 && v.inner > 1 && (align_a % 16) == 0 && !(flags & TTV_B200_FLAG_NO_VEC);
    if (forced == TTV_B200_KERNEL_COLX && !eligible) return TTV_B200_ERR_OPTS;
    const int mode = env_int("TTV_B200_USE_COLX", -1);                                 // -1 auto, 0 never, 1 whenever eligible
    const bool odd = (v.inner % Vx) != 0 || (align_c % 16) != 0;                       // the plain kernel would load narrow
    const bool pick = forced == TTV_B200_KERNEL_COLX ? true
                    : forced != 0 ? false
                    : mode == 1 ? eligible
                    : mode == 0 ? false
                    : (eligible && odd && v.inner * s >= 2048);
    if (pick) return choose_colx(s, dtype_is_complex(dtype) != 0, v, opts, sms, out);
  }

  l.threads = (uint32_t)env_int("TTV_B200_THREADS", 256);
  if (l.threads < 32 || l.threads > 256 || (l.threads % 32)) return TTV_B200_ERR_OPTS;
  // A row of one to three CTA-widths whose last width would be mostly empty (625 vector columns: 256 + 256 + 113) fits
  // CTAs of 128 threads better (4 x 128 + 113): complex<double> [15625, 25, 625] 6.19 -> 6.54 TB/s.
  if (env_int("TTV_B200_THREADS", 0) == 0 && v.inner > 1 && forced != TTV_B200_KERNEL_DOT) {
    const uint64_t v0 = (flags & TTV_B200_FLAG_NO_VEC) ? 1 : vmax_of(s);
    if (v.inner % v0 == 0) {
      const uint64_t cv0 = v.inner / v0, u256 = ceil_div(cv0, 256), u128 = ceil_div(cv0, 128);
      if (cv0 >= 256 && u256 <= 3 && cv0 * 100 < u256 * 256 * 85 && cv0 * 100 >= u128 * 128 * 95) l.threads = 128;
    }
  }
  const uint64_t NT = l.threads;
  const uint64_t vmax = (flags & TTV_B200_FLAG_NO_VEC) ? 1 : std::max<uint64_t>(1, 16 / s);

  uint64_t row_units = 0;      // COL with a row of 1..8 CTA-widths: the units it is spread over
  const bool dot = (v.inner == 1) && forced != TTV_B200_KERNEL_COL;
  if (forced == TTV_B200_KERNEL_DOT && v.inner != 1) return TTV_B200_ERR_OPTS;
  l.kernel = dot ? TTV_B200_KERNEL_DOT : TTV_B200_KERNEL_COL;

  int want = opts ? opts->ksplit : 0;
  if (want < 0) return TTV_B200_ERR_OPTS;
  if (want == 0) want = env_int("TTV_B200_KSPLIT", 0);

  uint64_t V = vmax;
  if (dot) {
    // vector along n_q: every fiber must start on a vector boundary and hold whole vectors ...
    while (V > 1 && !((v.nq % V) == 0 && (align_a % (V * s)) == 0 && (align_b % (V * s)) == 0)) V /= 2;
    // ... or, for fibers of odd length, be peeled into head | aligned body | tail (needs all of b in shared memory,
    // V shifted copies, and no n_q split)
    const uint64_t peel_smem = vmax * (ceil_div(v.nq, vmax) * vmax + vmax) * s;
    if (V < vmax && vmax > 1 && v.nq >= 4 * vmax && peel_smem <= 64 * 1024 && (align_a % 16) == 0 && want <= 1 &&
        v.outer >= sms * 8 && env_int("TTV_B200_PEEL", 1)) {
      V = vmax;
      l.peel = 1;
    }
    const uint64_t kv = v.nq / V;                          // vector steps per fiber
    l.tx = 1;
    // lanes per fiber: about eight vectors per lane, at most one warp (a fiber then reduces with shuffles only); four
    // per lane on tensors under 2 GiB, where a tile of two batches instead of four shortens the last wave of CTAs.
    // Measured on fibers of 256 / 512 floats (tools/sweep.py --set dotk, profiles/r01_dot_lanes_probe.txt): at
    // 512 MiB 6.54 / 6.53 against 6.38 / 6.37 TB/s, at 4 GiB 7.13 / 7.17 against 7.13 / 7.09, at 16 GiB 7.14 against 7.18.
    const bool small_tensor = v.outer * v.nq * s < (2ull << 30);
    uint64_t ty = std::min<uint64_t>(32, pow2_ceil(ceil_div(kv, (l.peel || !small_tensor) ? 8 : 4)));
    if (l.peel) ty = std::max<uint64_t>(ty, V == 4 ? 4 : 1);     // two rounds of ty lanes cover the 2V-2 head/tail elements
    // few fibers: put more lanes on each one, as long as every lane keeps at least four vectors
    while (!l.peel && ty < NT && kv / (ty * 2) >= 4 && ceil_div(v.outer, std::max<uint64_t>(1, NT / ty)) < sms * 2) ty *= 2;
    {
      const int ty_env = env_int("TTV_B200_DOT_TY", 0);                 // experiments: lanes per fiber (power of two)
      if (ty_env > 0 && !l.peel && pow2_ceil((uint64_t)ty_env) == (uint64_t)ty_env && (uint64_t)ty_env <= NT) ty = (uint64_t)ty_env;
    }
    const uint64_t to = std::max<uint64_t>(1, std::min<uint64_t>(NT / ty, v.outer));
    l.ty = (uint32_t)ty; l.to = (uint32_t)to;
    l.udir = 1;
    l.stream = ty >= 32 ? 1u : 0u;     // long fibers: every warp walks its own 100+ KB stream, keep it out of L1
  } else {
    // vector along inner: rows must hold whole vectors and start on vector boundaries
    while (V > 1 && !((v.inner % V) == 0 && (align_a % (V * s)) == 0 && (align_c % (V * s)) == 0)) V /= 2;
    const uint64_t cv = v.inner / V;                       // vector columns
    if (cv >= NT) {
      // a row is only a few CTA-widths long: the units of a thread must not outnumber the units of a row, or most of
      // its loads are predicated off (480 columns, 4 units of 256: measured [13200, 25, 480] c128 5.8 -> 6.7 TB/s with 2).
      // Spreading the row evenly over the units (2 x 240 instead of 256 + 224) measured no better, so tx stays the CTA.
      // (Measured again on 625 columns of complex<double>, 3 x 209 instead of 256 + 256 + 113: 6.06 against 6.19 TB/s;
      // CTAs of 128 threads do help there, see the top of this function.)
      const uint64_t U = ceil_div(cv, NT);
      l.tx = (uint32_t)(U <= 8 && env_int("TTV_B200_EVEN_UNITS", 0) ? ceil_div(cv, U) : NT);
      row_units = U <= 8 ? U : 0;
      l.ty = 1; l.to = 1; l.udir = 0;
    } else {
      l.tx = (uint32_t)cv;
      const uint64_t rem = NT / cv;
      uint64_t ty = 1;
      // short rows: lanes run along n_q too, so that a warp still reads one contiguous run of memory
      if (v.inner * s < 128 && v.nq >= 16) ty = std::min<uint64_t>(rem, pow2_floor(v.nq / 8));
      uint64_t to = std::max<uint64_t>(1, std::min<uint64_t>(rem / ty, v.outer));
      // few slabs: use the idle threads of the CTA along n_q
      while (ty * 2 <= rem / to && v.nq / (ty * 2) >= 8 && ceil_div(v.outer, to) < sms * 2) ty *= 2;
      l.ty = (uint32_t)ty; l.to = (uint32_t)to; l.udir = 1;
    }
    {
      // experiments: force the thread tile of the column kernel (tx lanes along inner, ty along n_q, to slabs per CTA)
      const int tx_env = env_int("TTV_B200_COL_TX", 0), ty_env = env_int("TTV_B200_COL_TY", 0), to_env = env_int("TTV_B200_COL_TO", 0);
      if (tx_env > 0 || ty_env > 0 || to_env > 0) {
        uint64_t tx = tx_env > 0 ? std::min<uint64_t>((uint64_t)tx_env, std::min(cv, NT)) : l.tx;
        uint64_t ty = ty_env > 0 ? std::min<uint64_t>((uint64_t)ty_env, std::max<uint64_t>(1, NT / tx)) : 1;
        ty = std::max<uint64_t>(1, std::min(ty, v.nq));
        uint64_t to = to_env > 0 ? (uint64_t)to_env : std::max<uint64_t>(1, NT / (tx * ty));
        to = std::max<uint64_t>(1, std::min(std::min(to, NT / (tx * ty)), v.outer));
        l.tx = (uint32_t)tx; l.ty = (uint32_t)ty; l.to = (uint32_t)to;
        l.udir = cv > tx ? 0u : 1u;
        row_units = 0;
        if (l.udir == 0) { const uint64_t U = ceil_div(cv, tx); row_units = U <= 8 ? U : 0; }
      }
    }
    // measured: 4/8-byte elements are faster through L1 unless lanes are strung along n_q; 16-byte elements bypass it
    l.stream = (s >= 16 || l.ty > 1) ? 1u : 0u;
  }
  l.stream = (uint32_t)env_int("TTV_B200_STREAM", (int)l.stream);
  l.vec = (int)V;

  // split n_q across CTAs only when there are too few tiles to occupy the SMs at all
  const uint64_t itiles1 = dot ? 1 : ceil_div(v.inner / V, l.tx);
  const uint64_t otiles1 = ceil_div(v.outer, l.to);
  const uint64_t tiles1  = itiles1 * otiles1;
  const uint64_t kstep   = (uint64_t)l.ty * (dot ? V : 1);            // n_q elements one pass of the CTA covers
  uint64_t ksplit = 1;
  if (want > 0) ksplit = (uint64_t)want;
  else if (tiles1 < sms && !l.peel) {
    const uint64_t max_split = std::max<uint64_t>(1, v.nq / (kstep * 16));
    ksplit = std::min(ceil_div(sms * 8, tiles1), max_split);
  }
  // Warp-per-fiber DOT on very long fibers: the resident warps all sit at the same offset of fibers that lie a large
  // power of two apart, which camps on a few DRAM partitions (65536^2 fp32, q=1: 6.9 TB/s; 7.3 TB/s when the fibers
  // are cut into ~8 KB pieces, because neighbouring CTAs then walk neighbouring pieces of the same fibers).
  if (want == 0 && dot && !l.peel && l.ty == 32 && ksplit == 1 && v.nq * s >= 64 * 1024)
    ksplit = std::min<uint64_t>(64, v.nq * s / 8192);
  if (l.peel) ksplit = 1;
  ksplit = std::max<uint64_t>(1, std::min(ksplit, ceil_div(v.nq, kstep)));
  uint64_t kchunk = ceil_div(ceil_div(v.nq, ksplit), kstep) * kstep;  // multiple of kstep keeps vectors aligned
  ksplit = ceil_div(v.nq, kchunk);
  l.ksplit = (uint32_t)ksplit;
  l.kchunk = kchunk;

  // Lanes strung along n_q consume b as fast as rows of A; when b is too long to stay resident, re-staging it chunk by
  // chunk puts two __syncthreads around every couple of batches (measured [1024, 262144, 4] fp32: 4.2 TB/s against
  // 7.0 TB/s).  Those lanes read consecutive elements of b anyway, so they take them straight from L2 inside the batch.
  {
    const int mode = env_int("TTV_B200_BDIRECT", -1);
    const bool can = l.ty > 1 && !l.peel && kchunk < (1ull << 31) && (!dot || (align_b % (V * s)) == 0);
    const bool resident = ksplit == 1 && v.nq * s <= 16384;
    if (can && (mode == 1 || (mode == -1 && !resident))) l.bdirect = 1;
  }

  // batch shape: ku k-steps for each of nu units; nu*ku loads in flight per thread (128 bytes with 16-byte vectors,
  // 16 loads with narrower ones)
  uint64_t loads = (V * s >= 16 || l.peel) ? 8 : 16;
  const uint64_t per = ceil_div(std::min(kchunk, v.nq), kstep);       // k-steps one thread makes per unit
  // Few CTAs (under ~10 per SM) cannot keep HBM busy with 128 bytes in flight per thread: the column kernel then runs
  // with 16 vector loads per batch (256 bytes, 2 CTAs of 128 registers per SM).  Measured on [1, n_q, 262144] fp32,
  // 256 CTAs: 6.38 -> 6.54 TB/s at 512 MB, 7.06 -> 7.31 at 4 GB; [1024, 512, 512], 512 CTAs: 5.96 -> 6.99; with 2048
  // CTAs and more the three-CTA form is ahead again, and so it is when n_q is split across CTAs (65536^2 q=2, 19
  // partitions: 7.34 against 7.05) (profiles/r01_small_launch_probe.txt).
  {
    const uint64_t units0 = dot ? ceil_div(v.outer, l.to) : (l.udir == 0 ? ceil_div(v.inner / V, l.tx) * ceil_div(v.outer, l.to) : ceil_div(v.outer, l.to));
    const int deep_env = env_int("TTV_B200_LOADS", 0);
    const bool few = !dot && ksplit == 1 && units0 < sms * 10 && per >= 16;       // (with n_q split across CTAs the 3-CTA form stays ahead)
    if (!l.peel && V * s >= 16 && (deep_env == 16 || (deep_env == 0 && few && !l.bdirect))) loads = 16;
  }
  // Batch depth (all rules measured on B200, tools/sweep.py).  Predicated-off slots of the last batch are wasted
  // issue slots, so short contractions take the depth that wastes least and fill the batch with more units; long ones
  // take the full depth with one unit.
  auto waste = [per](uint64_t d) { return (double)(ceil_div(per, d) * d - per) / (double)(ceil_div(per, d) * d); };
  auto least_waste = [&](uint64_t dmax, bool tie_deeper) {
    uint64_t best = 2;
    for (uint64_t d = 4; d <= dmax; d *= 2)
      if (waste(d) < waste(best) - 1e-9 || (tie_deeper && waste(d) <= waste(best) + 1e-9)) best = d;
    return best;
  };
  uint64_t ku;
  if (l.peel)                       ku = per >= 5 ? 8 : per >= 3 ? 4 : 2;       // per-unit head/tail work: go deep
  else if (dot && per <= 1)         ku = 1;
  else if (dot && per >= 64)        ku = loads;                                 // very long fibers: one fiber per lane group
  else if (dot && per >= 16)        ku = 4;
  else if (!dot && per >= 16) {                                                 // deepest batch that wastes <= 10 %
    ku = (loads > 8 && per < 64) ? 4 : loads;                                   // narrow loads: shallow, more units
    // (never below four: depth matters more than the predicated-off slots of the last batch -- 25 rows of
    // complex<double>: ku = 2 / 4 / 8 waste 4 / 11 / 22 % and run 5.3 / 6.7 / 6.4 TB/s)
    while (ku > 4 && waste(ku) > 0.10) ku /= 2;
  } else                            ku = least_waste(std::min<uint64_t>(loads, 8), false);
  const int ku_env = env_int("TTV_B200_KU", 0);
  if (ku_env > 0 && (uint64_t)ku_env <= loads && pow2_ceil((uint64_t)ku_env) == (uint64_t)ku_env && (ku_env > 1 || (dot && !l.peel))) ku = (uint64_t)ku_env;
  uint64_t nu = loads / ku;
  if (nu > 8) nu = 8;                                                 // (16,1) is not instantiated
  // units multiply the work of a tile: never starve the SMs of tiles for their sake
  {
    const uint64_t units1 = dot ? ceil_div(v.outer, l.to) : (l.udir == 0 ? ceil_div(v.inner / V, l.tx) * v.outer : ceil_div(v.outer, l.to));
    while (nu > 1 && ceil_div(units1, nu) * ksplit < sms * 8) {
      nu /= 2;
      ku *= 2;
    }
    if (!(nu == 8 && ku == 1)) ku = loads / nu;                      // only (nu, ku) with nu*ku == loads are instantiated
  }
  if (row_units) {                                                    // units per thread must divide the units of a row
    while (nu > 1 && row_units % nu) nu /= 2;
    ku = loads / nu;
  }
  if (l.bdirect) { ku = 8; nu = 1; }                                  // the one batch shape instantiated with direct b
  l.ku = (int)ku;
  l.nu = (int)nu;

  if (dot) {
    l.itiles = 1;
    l.otiles = ceil_div(v.outer, (uint64_t)l.to * l.nu);
    l.a_ustride = (uint64_t)l.to * v.nq;
    l.c_ustride = l.to;
  } else if (l.udir == 0) {
    l.itiles = ceil_div(v.inner / V, (uint64_t)l.tx * l.nu);
    l.otiles = ceil_div(v.outer, (uint64_t)l.to);
    l.a_ustride = (uint64_t)l.tx * V;
    l.c_ustride = (uint64_t)l.tx * V;
  } else {
    l.itiles = ceil_div(v.inner / V, (uint64_t)l.tx);
    l.otiles = ceil_div(v.outer, (uint64_t)l.to * l.nu);
    l.a_ustride = (uint64_t)l.to * v.nq * v.inner;
    l.c_ustride = (uint64_t)l.to * v.inner;
  }
  l.tiles = l.itiles * l.otiles * ksplit;
  l.ctas  = std::min<uint64_t>(l.tiles, sms * (uint64_t)std::max(1, env_int("TTV_B200_GRID_MULT", 64)));
  if (env_int("TTV_B200_GRID", 0) > 0) l.ctas = std::min<uint64_t>(l.tiles, (uint64_t)env_int("TTV_B200_GRID", 0));

  // shared memory: a chunk of b (16 KB at most; the peeled DOT keeps V shifted copies of all of b) + reduction scratch
  uint64_t kb = std::min<uint64_t>(kchunk, 16384 / s);
  kb = std::max<uint64_t>(kstep, kb / kstep * kstep);
  if (kb * s > 96 * 1024) return TTV_B200_ERR_OPTS;
  if (l.bdirect) kb = kchunk;
  l.kb = (uint32_t)kb;
  const uint64_t red_elems = NT * (uint64_t)l.nu * (dot ? 1 : V);
  const uint64_t b_bytes = l.peel ? V * (ceil_div(v.nq, V) * V + V) * s : l.bdirect ? 0 : kb * s;
  l.smem_bytes = b_bytes + red_elems * s;
  l.workspace_bytes = ksplit > 1 ? ksplit * v.outer * v.inner * s : 0;
  *out = l;
  return TTV_B200_OK;
}

void fill_plan(int dtype, const View& v, const Launch& l, ttv_b200_plan_t* plan)
{
  std::memset(plan, 0, sizeof *plan);
  const uint64_t s = (uint64_t)dtype_size(dtype);
  plan->outer = v.outer; plan->nq = v.nq; plan->inner = v.inner;
  plan->k = v.k; plan->ref_case = v.ref_case;
  plan->kernel = v.strided ? TTV_B200_KERNEL_STRIDED : l.kernel; plan->vec = v.strided ? 1 : l.vec; plan->tx = (int32_t)l.tx; plan->ty = (int32_t)l.ty; plan->to = (int32_t)l.to;
  plan->nu = l.nu; plan->ku = l.ku; plan->stream = (int32_t)l.stream;
  plan->ksplit = (int32_t)l.ksplit; plan->threads = (int32_t)l.threads; plan->ctas = l.ctas;
  plan->smem_bytes = l.smem_bytes;
  const uint64_t rest = v.outer * v.inner, total = rest * v.nq;
  plan->algo_bytes = s * (total + v.nq + rest);
  plan->algo_flops = (dtype_is_complex(dtype) ? 8 : 2) * total;
  plan->workspace_bytes = l.workspace_bytes;
}

} // namespace ttvb
