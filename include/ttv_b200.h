/*
 * ttv_b200.h -- C-ABI of the B200-native mode-q tensor-times-vector product  C = A x_q b.
 *
 * This is the drop-in boundary.  The reference (bassoy/ttv) has no FFI layer; its operator API is the
 * C-like template function
 *
 *     tlib::ttv::ttv(ep, sp, fp, q, p, a, na, wa, pia, b, nb, c, nc, wc, pic)      reference include/tlib/ttv.h:54-92
 *
 * and everything else (tensor-level ttv ttv.h:99-114, operator* ttv.h:122-127, ttvpy.ttv wrapped_ttv.cpp:18-78)
 * forwards to it.  Each ttv_b200_<type>() below replaces one instantiation of that template: same thirteen
 * arguments in the same order and with the same meaning (1-based modes, element strides, layout tuples),
 * plain pointers and sizes only.  `include/tlib/ttv.h` in this repo is the header-style C++17 host API on top
 * of these symbols; `ttv_b200/ttvpy.py` is the Python binding on top of the same symbols.
 *
 * Semantics
 *   - a, b, c may be HOST or DEVICE pointers (classified with cudaPointerGetAttributes).  Device pointers are
 *     used in place.  Host pointers are staged: H2D of A and b, kernel, D2H of C, all inside the call.
 *   - C is OVERWRITTEN (C = A x_q b), which is what every caller of the reference observes because all of them
 *     zero C first (tensor.h:70, wrapped_ttv.cpp:67, gtest_tlib_ttv.cpp:125) and what the BLAS build always does
 *     (beta = 0, matrix_times_vector.h:213-215).  TTV_B200_FLAG_ACCUMULATE selects C += A x_q b, the behaviour of
 *     the reference's non-BLAS column kernel (matrix_times_vector.h:124).
 *   - Argument checks, their order and their messages follow ttv.h:64-89 and tensor_times_vector.h:147-168.
 *   - The call is synchronous unless TTV_B200_FLAG_ASYNC is set together with device pointers.
 *   - Threads: every entry point may be called from several host threads, also on ONE stream (torch's default stream is
 *     shared by all threads of a process).  Device-pointer calls only share a small table (looked up under a short lock,
 *     never held across a launch or a wait); the partial sums of a split n_q are a stream-ordered allocation of the call
 *     itself, so interleaved launches never see each other's partials.  Host-pointer calls share the staging buffers of
 *     their device and therefore run one at a time per device (they are PCIe-bound).
 *     The reference is re-entrant apart from its process-global BLAS/OpenMP settings (tensor_times_vector.h:90-143).
 *   - There is no CPU fallback: without a usable CUDA device every compute entry returns TTV_B200_ERR_CUDA.
 *     ttv_b200_plan() is pure host code and needs no device.
 */
#ifndef TTV_B200_H
#define TTV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TTV_B200_VERSION 110

/* element types ----------------------------------------------------------------------------------------- */
enum ttv_b200_dtype {
  TTV_B200_F32  = 0,  /* float                                     */
  TTV_B200_F64  = 1,  /* double                                    */
  TTV_B200_C64  = 2,  /* std::complex<float>   (re, im) pairs      */
  TTV_B200_C128 = 3,  /* std::complex<double>  (re, im) pairs      */
  TTV_B200_I32  = 4,  /* int32_t, wrap-around arithmetic           */
  TTV_B200_I64  = 5,  /* int64_t, wrap-around arithmetic           */
  TTV_B200_DTYPE_COUNT = 6
};

/* status codes; 1..20 are the reference's throw sites in the order they are evaluated --------------------- */
enum ttv_b200_status {
  TTV_B200_OK = 0,
  TTV_B200_ERR_ORDER_ZERO      = 1,   /* ttv.h:64  */
  TTV_B200_ERR_MODE            = 2,   /* ttv.h:65  */
  TTV_B200_ERR_A_NULL          = 3,   /* ttv.h:66  */
  TTV_B200_ERR_B_NULL          = 4,   /* ttv.h:67  */
  TTV_B200_ERR_C_NULL          = 5,   /* ttv.h:68  */
  TTV_B200_ERR_NA_NULL         = 6,   /* ttv.h:70  */
  TTV_B200_ERR_NB_NULL         = 7,   /* ttv.h:71  */
  TTV_B200_ERR_NC_NULL         = 8,   /* ttv.h:72  */
  TTV_B200_ERR_WA_NULL         = 9,   /* ttv.h:74  */
  TTV_B200_ERR_WC_NULL         = 10,  /* ttv.h:75  */
  TTV_B200_ERR_PIA_NULL        = 11,  /* ttv.h:77  */
  TTV_B200_ERR_PIC_NULL        = 12,  /* ttv.h:78  */
  TTV_B200_ERR_EXTENT_MISMATCH = 13,  /* ttv.h:80  */
  TTV_B200_ERR_SHAPE_A         = 14,  /* ttv.h:82  */
  TTV_B200_ERR_SHAPE_C         = 15,  /* ttv.h:83  */
  TTV_B200_ERR_LAYOUT_A        = 16,  /* ttv.h:85  */
  TTV_B200_ERR_LAYOUT_C        = 17,  /* ttv.h:86  */
  TTV_B200_ERR_STRIDES_A       = 18,  /* ttv.h:88  */
  TTV_B200_ERR_STRIDES_C       = 19,  /* ttv.h:89  */
  TTV_B200_ERR_LAYOUT_BEGIN    = 20,  /* tensor_times_vector.h:160 */
  TTV_B200_ERR_LAYOUT_END      = 21,  /* tensor_times_vector.h:164 */
  /* not in the reference: conditions it leaves undefined or cannot meet */
  TTV_B200_ERR_NOT_PACKED      = 30,  /* non-packed strides whose free modes cannot be matched between A and C (nc is not
                                         na without q, or more than 8 unfoldable free modes).  Valid non-packed strides
                                         ARE honoured in case 8, like the reference's slice variants do
                                         (tensor_times_vector.h:189-216); cases 1-7 ignore wa/wc like mtv does */
  TTV_B200_ERR_DTYPE           = 31,
  TTV_B200_ERR_OPTS            = 32,
  TTV_B200_ERR_CUDA            = 40,  /* no device, launch or runtime failure; text in ttv_b200_last_error() */
  TTV_B200_ERR_MIXED_POINTERS  = 41   /* a, b, c must be all host or all device */
};

/* policy hints: mirror the reference's tag types (include/tlib/detail/tags.h:21-63).  All of them select the
 * same GPU path; they are accepted so that call sites keep compiling and recorded in the plan. */
enum ttv_b200_execution { TTV_B200_EXEC_SEQ = 0, TTV_B200_EXEC_SEQ_BLAS, TTV_B200_EXEC_PAR, TTV_B200_EXEC_PAR_LOOP,
                          TTV_B200_EXEC_PAR_TASKLOOP, TTV_B200_EXEC_PAR_TASK, TTV_B200_EXEC_PAR_BLAS,
                          TTV_B200_EXEC_PAR_BLAS_LOOP };
enum ttv_b200_slicing   { TTV_B200_SLICE = 0, TTV_B200_SUBTENSOR = 1 };
enum ttv_b200_fusion    { TTV_B200_FUSE_NONE = 0, TTV_B200_FUSE_OUTER = 1, TTV_B200_FUSE_ALL = 2 };

/* kernel families.  0 lets the chooser decide; the others force a family (used by the tests to
 * run every kernel on every shape it supports). */
enum ttv_b200_kernel {
  TTV_B200_KERNEL_AUTO   = 0,
  TTV_B200_KERNEL_DOT    = 1,  /* mode q contiguous (inner == 1): cooperating lanes per fiber, shuffle reduction   */
  TTV_B200_KERNEL_COL    = 2,  /* column GEMV: thread owns contiguous outputs, streams A along inner                */
  TTV_B200_KERNEL_STREAM = 3,  /* small inner / small n_q: slab staged through shared memory with bulk copies       */
  TTV_B200_KERNEL_COLX   = 4,  /* column GEMV for rows that start off 16-byte boundaries: phase lanes along n_q   */
  TTV_B200_KERNEL_DOTF   = 5,  /* mode q contiguous, short fibers: A read as one flat stream, partials via smem    */
  TTV_B200_KERNEL_STRIDED = 6, /* reported only: non-packed strides (case 8), general-stride kernel; cannot be forced  */
  TTV_B200_KERNEL_COLT   = 7,  /* column GEMV with A staged through shared memory by TMA tensor tiles (cp.async.bulk.tensor),
                                  producer warp + mbarrier ring; chosen only when forced (measured against COL) */
  TTV_B200_KERNEL_STREAMK = 8, /* tiny odd inner extent under a long contraction: rows and b staged through shared memory by
                                  bulk copies, threads along n_q, `inner` accumulators each, tree at the end; by rule only
                                  where COLF cannot go (several slabs that start off the 16-byte grid)                      */
  TTV_B200_KERNEL_DOTP   = 9,  /* fibers of two elements (inner = 1, n_q = 2; 4- and 8-byte types): a lane loads consecutive
                                  16-byte vectors of whole fibers and stores their 8 bytes of C, b in registers             */
  TTV_B200_KERNEL_COLF   = 10, /* rows narrower than / not a multiple of a 16-byte vector (inner = 2, 3, 5, 6 ...): V / gcd(inner, V)
                                  rows are whole vectors, so a warp streams its slab (or slab partition, or several short
                                  slabs side by side) flat, one accumulator per element of a lane's vector                   */
  TTV_B200_KERNEL_COUNT  = 11
};

enum ttv_b200_flags {
  TTV_B200_FLAG_ACCUMULATE = 1,   /* C += A x_q b                                                        */
  TTV_B200_FLAG_ASYNC      = 2,   /* device pointers only: return after enqueueing on opts->stream       */
  TTV_B200_FLAG_NO_VEC     = 4,   /* testing: force scalar loads                                          */
  TTV_B200_FLAG_HONOR_STRIDES = 8 /* take wa, wc and pic at their word in EVERY case: the reference (and this library
                                     by default) ignores them in cases 1-7 (mtv, matrix_times_vector.h:314-336).  With
                                     it, padded tensors and a C with a layout of its own work for any q; packed
                                     inputs still take the fast kernels.  The numpy / torch front ends set it for
                                     arrays that are not contiguous. */
};

typedef struct ttv_b200_opts {
  int32_t  device;      /* CUDA device ordinal, -1 = current device                                          */
  int32_t  execution;   /* enum ttv_b200_execution (hint)                                                    */
  int32_t  slicing;     /* enum ttv_b200_slicing   (hint)                                                    */
  int32_t  fusion;      /* enum ttv_b200_fusion    (hint)                                                    */
  int32_t  kernel;      /* enum ttv_b200_kernel, 0 = auto                                                    */
  int32_t  ksplit;      /* number of n_q partitions, 0 = auto, 1 = never split                               */
  uint32_t flags;       /* enum ttv_b200_flags                                                               */
  int32_t  reserved;
  void*    stream;      /* cudaStream_t; NULL = the legacy default stream                                    */
} ttv_b200_opts;

/* what the layout folder and the kernel chooser decided for one call -------------------------------------- */
typedef struct ttv_b200_plan_t {
  uint64_t outer;       /* product of the extents after  q in layout order                                   */
  uint64_t nq;          /* contraction extent na[q-1]                                                        */
  uint64_t inner;       /* product of the extents before q in layout order ( == wa[q-1] )                    */
  uint32_t k;           /* 1-based position of q in pia  ( pia^-1(q) )                                       */
  uint32_t ref_case;    /* the reference's case 1..8 (include/tlib/detail/cases.h:24-36)                     */
  int32_t  kernel;      /* enum ttv_b200_kernel chosen                                                       */
  int32_t  vec;         /* elements per vector load                                                          */
  int32_t  tx, ty, to;  /* thread tile: threads along inner / along n_q / along outer inside one CTA         */
  int32_t  nu, ku;      /* loads in flight per thread: ku k-steps for each of nu independent outputs          */
                        /* (COLF: tx = vectors per super-row, ty = super-rows per step of a slab's lane group, */
                        /*  to = rows per super-row, nu = slabs a warp works on side by side)                  */
  int32_t  ksplit;      /* n_q partitions (second pass reduces them when > 1)                                */
  int32_t  threads;     /* threads per CTA                                                                   */
  int32_t  stream;      /* 1: A is loaded with L1::no_allocate                                               */
  uint64_t ctas;        /* grid size                                                                         */
  uint64_t smem_bytes;  /* dynamic shared memory per CTA                                                     */
  uint64_t algo_bytes;  /* sizeof(T) * (N + n_q + N/n_q): read A once, read b once, write C once             */
  uint64_t algo_flops;  /* 2*N (real) or 8*N (complex)                                                       */
  uint64_t workspace_bytes;
} ttv_b200_plan_t;

/* generic entry; the typed entries below forward to it */
int ttv_b200_run(int dtype, uint64_t q, uint64_t p,
                 const void* a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,
                 const void* b, const uint64_t* nb,
                 void* c, const uint64_t* nc, const uint64_t* wc, const uint64_t* pic,
                 const ttv_b200_opts* opts);

/* one entry per element type: replaces tlib::ttv::ttv<value_t,size_t,...>  (ttv.h:54-92) */
int ttv_b200_f32 (uint64_t q, uint64_t p, const float*   a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,
                  const float*   b, const uint64_t* nb, float*   c, const uint64_t* nc, const uint64_t* wc,
                  const uint64_t* pic, const ttv_b200_opts* opts);
int ttv_b200_f64 (uint64_t q, uint64_t p, const double*  a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,
                  const double*  b, const uint64_t* nb, double*  c, const uint64_t* nc, const uint64_t* wc,
                  const uint64_t* pic, const ttv_b200_opts* opts);
int ttv_b200_c64 (uint64_t q, uint64_t p, const void*    a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,
                  const void*    b, const uint64_t* nb, void*    c, const uint64_t* nc, const uint64_t* wc,
                  const uint64_t* pic, const ttv_b200_opts* opts);
int ttv_b200_c128(uint64_t q, uint64_t p, const void*    a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,
                  const void*    b, const uint64_t* nb, void*    c, const uint64_t* nc, const uint64_t* wc,
                  const uint64_t* pic, const ttv_b200_opts* opts);
int ttv_b200_i32 (uint64_t q, uint64_t p, const int32_t* a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,
                  const int32_t* b, const uint64_t* nb, int32_t* c, const uint64_t* nc, const uint64_t* wc,
                  const uint64_t* pic, const ttv_b200_opts* opts);
int ttv_b200_i64 (uint64_t q, uint64_t p, const int64_t* a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,
                  const int64_t* b, const uint64_t* nb, int64_t* c, const uint64_t* nc, const uint64_t* wc,
                  const uint64_t* pic, const ttv_b200_opts* opts);

/* Several products on ONE tensor:  C_i = A x_{q[i]} b_i,  i < count  (the reference's benchmark protocol contracts every
 * mode q = 1..p of the same tensor, README.md:59-64).  Equivalent to `count` calls of ttv_b200_run with nb = na[q_i-1]
 * and C_i packed in the output shape / layout of (na, pia, q_i) (detail/shape.h:126-158, detail/layout.h:175-207) --
 * except that with HOST pointers A crosses PCIe once instead of `count` times.  a, b[i], c[i]: all host or all device. */
int ttv_b200_multi(int dtype, uint64_t p,
                   const void* a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,
                   uint64_t count, const uint64_t* q, const void* const* b, void* const* c,
                   const ttv_b200_opts* opts);

/* The chain of p-1 products that leaves mode q:  c = A x_1 b_1 ... x_{q-1} b_{q-1} x_{q+1} b_{q+1} ... x_p b_p, a vector of
 * na[q-1] elements.  Replaces ttvpy::ttvs (reference ttvpy/src/wrapped_ttv.cpp:83-198), which calls ttv p-1 times and
 * creates a fresh numpy array for every intermediate.  Here the intermediates live in HBM (stream-ordered pool of the
 * library): with HOST pointers A crosses PCIe once, streamed in chunks under the kernels of the first product, the
 * other p-2 kernels run back to back on the device and only c comes back; with DEVICE pointers nothing leaves HBM and
 * TTV_B200_FLAG_ASYNC returns after enqueueing on opts->stream.
 *   a        packed tensor of shape na and layout pia (ttvpy: C-contiguous numpy array = last-order layout p..1)
 *   b[j]     j < p-1, the vector of mode j+1 (j+1 < q) or j+2: na[mode-1] elements, unit stride (wrapped_ttv.cpp:126-129)
 *   order    which mode goes first (wrapped_ttv.cpp:135-192): longest vector first / mode p downwards / mode 1 upwards;
 *            every order gives the same result up to the rounding of a different summation order
 *   a, b[j], c: all host or all device.  opts: device, stream, hints and TTV_B200_FLAG_{ASYNC,NO_VEC}; other fields ignored. */
enum ttv_b200_chain_order { TTV_B200_CHAIN_OPTIMAL = 0, TTV_B200_CHAIN_BACKWARD = 1, TTV_B200_CHAIN_FORWARD = 2 };
int ttv_b200_ttvs(int dtype, uint64_t q, uint64_t p, const void* a, const uint64_t* na, const uint64_t* pia,
                  const void* const* b, int order, void* c, const ttv_b200_opts* opts);
/* The schedule ttv_b200_ttvs follows, p-1 entries: step i contracts mode modes[i] (1-based, numbered in the tensor left
 * after the steps before it) with vector b[vectors[i]].  Pure host code. */
int ttv_b200_chain_plan(uint64_t q, uint64_t p, const uint64_t* na, int order, uint64_t* modes, uint64_t* vectors);

/* Validation + layout folding + kernel choice without touching a device (pure host code).  Pointers a, b, c are
 * only checked for null-ness; pass any non-null value.  Returns the same status codes as ttv_b200_run. */
int ttv_b200_plan(int dtype, uint64_t q, uint64_t p,
                  const void* a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,
                  const void* b, const uint64_t* nb,
                  const void* c, const uint64_t* nc, const uint64_t* wc, const uint64_t* pic,
                  const ttv_b200_opts* opts, ttv_b200_plan_t* plan);

/* The canonical view directly: C[outer][inner] (=|+=) sum_k A[outer][k][inner] * b[k], everything packed, DEVICE
 * pointers only.  This is the shape every legal (na, pia, q) folds to; the sharded multi-GPU driver and the
 * detail::gemv_* shims (reference matrix_times_vector.h:51-127) call it. */
int ttv_b200_view(int dtype, uint64_t outer, uint64_t nq, uint64_t inner,
                  const void* a, const void* b, void* c, const ttv_b200_opts* opts);
int ttv_b200_plan_view(int dtype, uint64_t outer, uint64_t nq, uint64_t inner,
                       const ttv_b200_opts* opts, ttv_b200_plan_t* plan);

/* Multi-GPU, n_q split (mode q is the slowest mode; every GPU holds a range of the contraction rows): the product of the
 * canonical view FUSED with its exchange step.  The flat index space of C (outer*inner elements) is cut into `world`
 * blocks of `blk` elements (blk a multiple of 16 bytes' worth of elements, world*blk >= outer*inner); this GPU's partial
 * sums for block j are written by the kernel's own stores into peer_ws[j] + rank*blk, where peer_ws[j] is GPU j's
 * workspace of world*blk elements mapped into this process (NVLink peer memory, e.g. torch symmetric memory).  After a
 * barrier across the GPUs, ttv_b200_reduce_slots on every GPU sums the `world` slots of its own workspace in rank order
 * into its block of C (n <= blk elements): a deterministic reduce-scatter.  DEVICE pointers only; a, b, peer_ws[rank]
 * live on the current device.  Replaces the ncclReduce of SURVEY 8(e) for this case. */
int ttv_b200_view_scatter(int dtype, uint64_t outer, uint64_t nq, uint64_t inner, const void* a, const void* b,
                          void* const* peer_ws, uint32_t world, uint32_t rank, uint64_t blk, const ttv_b200_opts* opts);
int ttv_b200_reduce_slots(int dtype, const void* ws, void* c, uint64_t n, uint64_t blk, uint32_t slots,
                          const ttv_b200_opts* opts);
/* The same exchange in ONE kernel launch per GPU: product + scatter as above, then -- inside the kernel -- a barrier across the
 * GPUs (the last CTA of a GPU to finish writes `token` into flag[rank] of every GPU's flag array over NVLink; every CTA waits
 * with ld.acquire.sys until all `world` flags of its own array show the token) and the sum of the `world` slots of this GPU's
 * workspace, in rank order, into c_block (n_block <= blk elements: this GPU's block of the flat C).  No library barrier, no
 * second and third launch.
 *   peer_flags[j]  GPU j's flag array (>= 16 x uint32, zero before the first round), mapped like peer_ws
 *   token          round number of the group: 1, 2, 3, ... (every GPU passes the same value; flags only grow)
 *   scratch        16 bytes of LOCAL device memory, zeroed once: arrival counter (8) + error flag (4)
 *   max_ctas       how many of the LAST CTAs to finish stay for the barrier and the slot sum; 0 = one per SM.  Only they
 *                  wait, so the grid is the usual oversubscribed one.  Peer workspaces must alternate between two halves
 *                  from round to round.
 * Asynchronous with TTV_B200_FLAG_ASYNC; a wait longer than 10 s (TTV_B200_EXCHANGE_TIMEOUT_MS) sets the error flag, which
 * synchronous calls report as TTV_B200_ERR_CUDA. */
int ttv_b200_view_exchange(int dtype, uint64_t outer, uint64_t nq, uint64_t inner, const void* a, const void* b,
                           void* const* peer_ws, void* const* peer_flags, uint32_t world, uint32_t rank, uint64_t blk,
                           void* c_block, uint64_t n_block, uint32_t token, void* scratch, uint32_t max_ctas,
                           const ttv_b200_opts* opts);

/* x[i] = synth(seed, first + i), i < count, written on the device by a kernel (x is a DEVICE pointer).  The generator
 * is the counter-based splitmix64 one of SURVEY 8(d); oracle/ttv_oracle.c carries the identical host version, so
 * tensors too large for host memory can be checked by sampling. */
int ttv_b200_fill(int dtype, void* x, uint64_t first, uint64_t count, uint64_t seed, const ttv_b200_opts* opts);

/* ---- device / pinned memory for header-style hosts that do not include the CUDA runtime --------------------------------
 * tlib::ttv::device_tensor (include/tlib/detail/device_tensor.h) is built on these.  ttv_b200_copy moves bytes in any
 * direction (host <-> device, device <-> device); large transfers from or to PAGEABLE host memory are pipelined through the
 * library's pinned bounce buffers and copy threads, like the host-pointer path of ttv_b200_run.  Synchronous. */
int ttv_b200_device_alloc(void** ptr, uint64_t bytes, int device /* -1 = current */, int zero);
int ttv_b200_device_free(void* ptr);
int ttv_b200_host_alloc(void** ptr, uint64_t bytes);      /* page-locked host memory: H2D / D2H by DMA straight from it */
int ttv_b200_host_free(void* ptr);
int ttv_b200_copy(void* dst, const void* src, uint64_t bytes, const ttv_b200_opts* opts);

/* ---- a HOST tensor that keeps its copy in HBM between products -----------------------------------------------------------
 * The reference's benchmark protocol contracts every mode q = 1..p of one tensor (README.md:59-64), and its tensor class is
 * a std::vector in host memory (detail/tensor.h:56-114): called as a drop-in, every `A(q) * b` would move all of A across
 * PCIe again.  A ttv_b200_resident is the device-side twin of ONE host tensor: the first product after creation /
 * invalidation streams A across PCIe in chunks under its own kernels (as ttv_b200_run does) but INTO a buffer that stays;
 * every later product with the same (a, bytes) reads HBM and only b and C cross the bus.  The caller says when the host data
 * changed (ttv_b200_resident_invalidate); tlib::ttv::tensor does so from its mutating accessors when
 * tensor::keep_on_device(true) was asked for.  a, b, c are HOST pointers.  Same checks, semantics and status codes as
 * ttv_b200_run.  One resident object serves one thread at a time. */
typedef struct ttv_b200_resident ttv_b200_resident;
int  ttv_b200_resident_create(ttv_b200_resident** r, int device /* -1 = current */);
void ttv_b200_resident_destroy(ttv_b200_resident* r);
void ttv_b200_resident_invalidate(ttv_b200_resident* r);
int  ttv_b200_resident_valid(const ttv_b200_resident* r);    /* 1: the next product of the same tensor will not upload A */
int  ttv_b200_run_resident(ttv_b200_resident* r, int dtype, uint64_t q, uint64_t p,
                           const void* a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,
                           const void* b, const uint64_t* nb,
                           void* c, const uint64_t* nc, const uint64_t* wc, const uint64_t* pic,
                           const ttv_b200_opts* opts);

/* ---- one HOST tensor over several GPUs of this process ------------------------------------------------------------------
 * ttv_b200_run with host pointers, the work cut along the slowest mode of A's layout over `n_devices` GPUs, one host thread
 * per GPU, every GPU pulling its slab over its OWN PCIe link (SURVEY 8b "device list", 8e): q not the slowest mode -> free
 * split, every GPU returns its slab of C; q the slowest mode -> n_q split, the partial sums meet on devices[0] (peer copies)
 * and are added there in device order by ttv_reduce_kernel (deterministic), then C goes back.  Pays off for PINNED host
 * tensors (ttv_b200_host_alloc, cudaHostRegister, torch pin_memory): pageable memory is bounced through copy threads, whose
 * memcpy rate one GPU's link already matches.  devices may repeat (tests on a single-GPU box).  Small or strided inputs run
 * on devices[0] alone. */
int ttv_b200_run_devices(int dtype, uint64_t q, uint64_t p,
                         const void* a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,
                         const void* b, const uint64_t* nb,
                         void* c, const uint64_t* nc, const uint64_t* wc, const uint64_t* pic,
                         const ttv_b200_opts* opts, const int32_t* devices, uint32_t n_devices);

/* L0 helpers of the reference, restated (shape.h, layout.h, strides.h); pure host code ------------------- */
int ttv_b200_is_valid_shape  (const uint64_t* n,  uint64_t p);                       /* shape.h:30-34    */
int ttv_b200_is_valid_layout (const uint64_t* pi, uint64_t p);                       /* layout.h:29-55   */
int ttv_b200_is_valid_strides(const uint64_t* pi, uint64_t p, const uint64_t* w);    /* strides.h:76-101 */
int ttv_b200_compute_strides (const uint64_t* n,  const uint64_t* pi, uint64_t p, uint64_t* w);    /* strides.h:31-57  */
int ttv_b200_output_shape    (const uint64_t* na, uint64_t p, uint64_t q, uint64_t* nc);           /* shape.h:103-123  */
int ttv_b200_output_layout   (const uint64_t* pia, uint64_t p, uint64_t q, uint64_t* pic);         /* layout.h:143-172 */
int ttv_b200_k_order_layout  (uint64_t p, uint64_t k, uint64_t* pi);                               /* layout.h:57-76   */

/* diagnostics */
const char* ttv_b200_strerror(int status);     /* the reference's message text for codes 1..21 */
const char* ttv_b200_last_error(void);         /* thread-local; message of the last failing call on this thread */
int         ttv_b200_version(void);
int         ttv_b200_device_count(void);       /* 0 when no usable CUDA device */
uint64_t    ttv_b200_launch_count(void);       /* kernels launched by this library in this process */
int         ttv_b200_dtype_size(int dtype);
void        ttv_b200_release(void);            /* waits for queued work, frees workspaces / staging buffers / copy stream;
                                                  call it while no other entry point is running on another thread */

#ifdef __cplusplus
}
#endif
#endif /* TTV_B200_H */
