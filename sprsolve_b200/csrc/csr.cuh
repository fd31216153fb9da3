// csr.cuh -- device-resident CSR operator: the drop-in for MklMat<T> / CsMatI<T,i32>
// (src/mkl_mat.rs:15-22, src/mat.rs:47-183).
#pragma once
#include <vector>

#include "common.cuh"
#include "peer.cuh"

namespace spb {

// What a SpMV launch folds into its epilogue (each replaces a separate vecalg::conj_dot pass).
enum SpmvEpiMode {
  EPI_NONE = 0,
  EPI_DOT_WY = 1,  // slot0 = sum conj(w_i) * y_i            (mul_vec_dot: w = v_in; <r0,v>: w = r0)
  EPI_TT_TR = 2    // slot0 = sum conj(y_i) * y_i, slot1 = sum conj(y_i) * w_i   (<t,t>, <t,r>)
};

static const int kMaxXwinSegs = 32;

struct HaloPeer {
  int rank;
  int64_t send_off, send_cnt;  // into sendbuf / send_idx
  int64_t recv_off, recv_cnt;  // into halo
};

struct PeerWindow;  // dist.cuh

// Where this rank's boundary entries land in the neighbours' halo windows (peer transport).
struct PutArgs {
  int npeers;
  void* dst0[kMaxPeers];                 // peer's halo payload + my offset in it (parity 0), T*
  long long dst_stride[kMaxPeers];       // peer's n_halo: parity 1 lives dst_stride elements further
  unsigned long long* rflag[kMaxPeers];  // &peer_head->flags[my rank]
  unsigned long long* rack[kMaxPeers];   // &peer_head->acks[my rank]
  long long send_off[kMaxPeers + 1];     // into send_idx
};

template <typename T>
struct CsrMat : spb_op {
  int64_t nnz = 0;
  bool ip64 = false;
  DevBuf indptr;  // int32 or int64 [n_local+1]
  DevBuf cols;    // int32 [nnz + pad], LOCAL column ids: [0,n_local) owned, [n_local, n_local+n_halo) halo
  DevBuf vals;    // T [nnz + pad]
  int64_t max_row = 0;

  // --- analysis (the mkl_sparse_optimize analogue): nnz-balanced row tiles -----------------
  int plan_ct = 128;     // consumer threads per CTA (one row per thread and tile)
  int plan_stages = 2;   // shared-memory stages of the bulk-copy ring
  int plan_gb = 0;       // reserved
  int plan_tile = 2048;  // non-zeros staged per tile
  int plan_rcap = 264;   // indptr entries staged per tile
  int plan_bps = 1;      // resident CTAs per SM (occupancy)
  int64_t span = 0;      // tile t = rows whose first nnz lies in [t*span, (t+1)*span)
  int64_t ntiles = 0;
  DevBuf tile_row;       // int32 [ntiles+1]
  DevBuf partials;       // T [2 * max grid], per-block epilogue partial sums
  DevBuf red;            // scal2[2]: finalised epilogue sums

  // --- column-offset dictionary (stencil-like matrices; spmv.cu) ---------------------------
  bool dict_on = false;
  int dict_w = 0;        // dictionary row stride = longest row
  int64_t dict_u = 0;    // number of distinct row patterns
  DevBuf dict_off;       // int32 [dict_u * dict_w]: col - row of the pattern's entries
  DevBuf pid;            // uint32 [n_local + 1]: pattern id | (low 16 bits of indptr[row]) << 16 (the DICT kernels' row words)
  // --- x window (spmv.cu): the distinct column offsets of a stencil-like matrix form a few runs of
  //     consecutive values; per tile the x entries of every run are ONE contiguous segment, staged in
  //     shared memory with a bulk copy, and the gathers become shared-memory reads
  bool xwin_on = false;
  bool xwin_want = false;            // SPB_SPMV_XWIN=1, or chosen by autotune()
  std::vector<int> dict_off_host, dict_len_host;  // host copies of the dictionary (analysis of the runs)
  int xwin_nseg = 0;                 // runs of consecutive offsets
  int xwin_rows = 0;                 // most rows of a tile the window is laid out for
  int xwin_elems = 0;                // shared-memory elements of one window
  int xwin_gmin[kMaxXwinSegs] = {};  // first offset of the run
  int xwin_pad[kMaxXwinSegs] = {};   // elements the copy starts before it (16-byte alignment)
  int xwin_extra[kMaxXwinSegs] = {}; // pad + (last - first offset): copy length = extra + rows, rounded up
  int xwin_start[kMaxXwinSegs] = {}; // first shared-memory element of the segment
  DevBuf dict_soff;                  // int32 [dict_u * dict_w]: window element of (local row 0, entry)

  // --- row-block partition (multi-GPU) ------------------------------------------------------
  int64_t n_halo = 0;
  DevBuf halo;               // T [n_halo]: remote x entries, grouped by owner rank
  DevBuf sendbuf;            // T [total send]
  DevBuf send_idx;           // int32 [total send]: local indices to pack
  std::vector<HaloPeer> peers;
  std::vector<int32_t> halo_cols_global;  // sorted global ids of the halo slots (host copy)
  DevBuf tiles_interior, tiles_boundary;  // int32 tile lists (boundary = touches a halo column)
  DevBuf tiles_all;                       // interior tiles followed by boundary tiles (fused launch)
  int64_t n_tiles_interior = 0, n_tiles_boundary = 0;
  // peer transport: the halo lives in a window the neighbours write into (2 x n_halo, parity)
  PeerWindow* halo_win = nullptr;
  PutArgs put;
  ~CsrMat();

  // --- host-slice staging (trait methods on &[T]) --------------------------------------------
  DevBuf stage_in, stage_out;

  void analyze();
  void choose_format();
  void static_plan(int& consumer_threads, int& stages);
  void build_plan(int consumer_threads, int stages);
  void autotune();
  // y = A x (x, y device pointers to LOCAL vectors).  conj_in: multiply by conj(x) (CSMinRes,
  // src/cs_minres.rs:99-101, without materialising tvec).  Epilogue sums land in `red` after
  // finalize_epilogue() (local sum only; the caller all-reduces when distributed).
  void mul(const T* x, T* y, int epi_mode, const T* w, bool conj_in);
  int64_t last_partial_blocks = 0;
  // Sums the per-block partials of the last mul() into red[0], red[1] (device).  allreduce: also
  // sum over ranks (fused into the same kernel on the peer transport, ncclAllReduce otherwise).
  void finalize_epilogue(bool allreduce = false);
};

template <typename T>
CsrMat<T>* csr_from_host(Ctx* ctx, int64_t n_global, int64_t row_begin, int64_t row_end,
                         const void* indptr, int indptr_bits, const int32_t* indices,
                         const void* values);
template <typename T>
CsrMat<T>* csr_from_stencil(Ctx* ctx, int kind, int64_t nx, int64_t ny, int64_t nz,
                            const double* params, int nparams);

// ingest.cu: the other layouts callers start from (single GPU).  All arrays are host pointers.
template <typename T>
CsrMat<T>* csr_from_triplets(Ctx* ctx, int64_t n, int64_t nnz, const int32_t* rows, const int32_t* cols, const void* vals);
template <typename T>
CsrMat<T>* csr_from_csc(Ctx* ctx, int64_t n, const void* indptr, int indptr_bits, const int32_t* row_indices, const void* vals);
template <typename T>
CsrMat<T>* csr_from_matrix_market(Ctx* ctx, const char* path);

// dist.cu: turn GLOBAL column ids into local + halo ids, build the exchange plan.
template <typename T>
void csr_localize(CsrMat<T>* m);
template <typename T>
void classify_tiles(CsrMat<T>* m);  // interior / boundary tile lists (after analyze's tiling)
template <typename T>
void halo_exchange_begin(CsrMat<T>* m, const T* x);  // pack + send/recv on the comm stream
template <typename T>
void halo_exchange_wait(CsrMat<T>* m);               // compute stream waits for the halo
template <typename T>
void halo_put(CsrMat<T>* m, const T* x);             // peer transport: store into the neighbours' windows
template <typename T>
void halo_release(CsrMat<T>* m);                     // frees the halo window (collective-free)

}  // namespace spb
