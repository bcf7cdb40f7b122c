// plan.h -- argument validation, the stride/layout folder and the kernel chooser.  Pure host code, no CUDA.
//
// Replaces (reference include/tlib/): the sixteen checks of ttv.h:64-89, the 8-case classifier detail/cases.h:24-36,
// compute_inverse_pia_m / compute_ninvpia (detail/tensor_times_vector.h:147-180) and the whole policy dispatch of
// detail/tensor_times_vector.h:430-1361.  Every legal (na, pia, q) collapses to ONE canonical packed view
//
//        A[outer][nq][inner]      b[nq]      C[outer][inner]            (inner fastest)
//
// with k = pia^-1(q), inner = prod_{r<k} na[pia_r], outer = prod_{r>k} na[pia_r].  The reference's cases are the
// special values of that view: cases 2,5,6 (row GEMV) are inner == 1, cases 3,4,7 (column GEMV) are outer == 1,
// case 8 (the loop nest) is everything else.
#pragma once

#include <cstdint>
#include "../../include/ttv_b200.h"

namespace ttvb {

constexpr int kMaxOrder = 64;

constexpr int kMaxFreeDims = 8;

struct View {
  uint64_t outer = 1, nq = 1, inner = 1;
  uint32_t k = 0;          // 1-based position of q in pia
  uint32_t ref_case = 0;   // 1..8
  // the slowest mode of A's layout with extent > 1: along it A (and C, or n_q when it is mode q) splits into contiguous
  // slabs -- what the host-pointer path streams chunk by chunk.  0 = derive from the view (outer, else n_q).
  uint64_t slow_extent = 0;
  // General strides (case 8 only, like the reference's slice variants): the free modes in the order of C's layout,
  // fastest first, neighbours that are packed against each other in both tensors folded into one.
  bool     strided = false;
  uint32_t nfree = 0;
  uint64_t fn[kMaxFreeDims] = {0}, fwa[kMaxFreeDims] = {0}, fwc[kMaxFreeDims] = {0};
  uint64_t wq = 0;                     // stride of mode q in A
  uint64_t span_a = 0, span_c = 0;     // elements from the first to one past the last element touched
};

// How one launch of the tile kernel is shaped.  See kernels.cuh for the meaning of the thread tile.
struct Launch {
  int      kernel  = TTV_B200_KERNEL_COL;
  int      vec     = 1;     // elements per vector load (along inner for COL, along n_q for DOT)
  int      ku      = 8;     // k-steps of one unit in flight per thread
  int      nu      = 1;     // independent units (outputs) per thread; nu*ku vector loads are in flight per thread
  uint32_t udir    = 0;     // units run along inner (0) or along outer (1)
  uint32_t stream  = 1;     // L1::no_allocate loads
  uint32_t peel    = 0;     // DOT only: n_q is not a multiple of vec; fibers are split into head | aligned body | tail
  uint64_t a_ustride = 0, c_ustride = 0;   // element distance between two units in A / C
  uint32_t tx = 1, ty = 1, to = 1;   // threads along inner / along n_q / along outer inside one CTA
  uint32_t threads = 256;
  uint32_t ksplit  = 1;     // n_q partitions across CTAs; > 1 => partials in the workspace + reduce pass
  uint64_t kchunk  = 0;     // n_q elements per partition
  uint64_t itiles  = 1;     // tiles along inner
  uint64_t otiles  = 1;     // tiles along outer
  uint64_t tiles   = 1;     // itiles * otiles * ksplit
  uint64_t ctas    = 1;     // grid size (<= tiles; CTAs stride over tiles)
  uint32_t kb      = 0;     // elements of b staged in shared memory per step
  uint64_t smem_bytes = 0;
  uint64_t workspace_bytes = 0;
  // STREAM kernel only: slabs per shared-memory stage, bytes of one stage, number of chunks
  uint64_t slabs_per_chunk = 0, chunks = 0;
  uint32_t stage_bytes = 0;
  uint32_t stages = 3;      // STREAM: shared-memory stages in the ring (3 small ones x 2 CTAs per SM, or 3..5 of one big slab each)
  // COLX kernel only: output columns owned by one tile
  uint64_t wcols = 0;
  uint32_t bdirect = 0;     // b read straight from global memory / L2 (lanes along n_q, b too long to stay resident)
  uint32_t warp = 0;        // COLX: warp-autonomous form (ttv_colw_kernel)
  uint32_t pair = 0;        // COLF: rows of two 4-byte elements, b and C 8-byte aligned (ttv_colf2_kernel)
  uint32_t short1 = 0;      // COLF: a slab is at most one batch of its lane group, unsplit (ttv_colfs_kernel)
  uint32_t tiny = 0;        // COLF: slabs of 1 .. 16 vectors of two-element rows, transposing butterfly (ttv_colf_tiny_kernel)
  // STREAMK kernel only: rows per shared-memory stage (slabs_per_chunk), bytes of a stage of rows (stage_bytes) / of b
  uint32_t b_stage_bytes = 0;
  // COLT kernel only: box of the TMA tensor tile (wt 32-bit words of a row x kt rows), boxes per column tile
  uint32_t wt = 0, kt = 0, kboxes = 0;
};

int dtype_size(int dtype);          // bytes, 0 if unknown
int dtype_is_complex(int dtype);

// L0 helpers (restated from detail/shape.h, layout.h, strides.h)
bool is_valid_shape(const uint64_t* n, uint64_t p);
bool is_valid_layout(const uint64_t* pi, uint64_t p);
bool is_valid_strides(const uint64_t* pi, uint64_t p, const uint64_t* w);
int  compute_strides(const uint64_t* n, const uint64_t* pi, uint64_t p, uint64_t* w);
int  output_shape(const uint64_t* na, uint64_t p, uint64_t q, uint64_t* nc);
int  output_layout(const uint64_t* pia, uint64_t p, uint64_t q, uint64_t* pic);
int  k_order_layout(uint64_t p, uint64_t k, uint64_t* pi);
int  classify_case(uint64_t p, uint64_t q, const uint64_t* pia);

// Validation in the reference's order, then folding.  Returns a ttv_b200_status.
int validate_and_fold(uint64_t q, uint64_t p,
                      const void* a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,
                      const void* b, const uint64_t* nb,
                      const void* c, const uint64_t* nc, const uint64_t* wc, const uint64_t* pic,
                      uint32_t flags, View* view);

// Kernel choice for a canonical view.  align_a / align_c are the byte alignments of the device pointers (use 256
// when unknown, e.g. in ttv_b200_plan).  sm_count = number of SMs of the target device (148 on B200).
int choose_launch(int dtype, const View& v, const ttv_b200_opts* opts, uint64_t align_a, uint64_t align_b,
                  uint64_t align_c, int sm_count, Launch* out);

void fill_plan(int dtype, const View& v, const Launch& l, ttv_b200_plan_t* plan);

const char* status_message(int status);

} // namespace ttvb
