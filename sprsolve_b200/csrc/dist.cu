// dist.cu -- multi-GPU support: the matrix is partitioned by contiguous row blocks, one process
// (one spb_ctx) per GPU.  Per SpMV the x entries of off-block columns ("halo") travel over
// NVLink with grouped ncclSend/ncclRecv on a separate stream while the interior rows are being
// multiplied; the Krylov scalars are summed with ncclAllReduce.  The reference has no analogue
// (single process; rayon / MKL threads only, src/mat.rs:85-107).
#include <cub/cub.cuh>
#include <dlfcn.h>

#include <string.h>

#include <algorithm>
#include <mutex>

#include "csr.cuh"
#include "dist.cuh"

namespace spb {

// ---------------------------------------------------------------- NCCL via dlopen
static NcclApi g_nccl;
static bool g_nccl_ok = false;
static std::once_flag g_nccl_once;

static void load_nccl() {
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // torch's copy, if loaded
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return;
  bool ok = true;
  auto sym = [&](const char* n, bool required = true) {
    void* p = dlsym(h, n);
    if (!p && required) ok = false;
    return p;
  };
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
  g_nccl.CommSplit = (decltype(g_nccl.CommSplit))sym("ncclCommSplit", false);
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
  g_nccl.AllGather = (decltype(g_nccl.AllGather))sym("ncclAllGather");
  g_nccl.Send = (decltype(g_nccl.Send))sym("ncclSend");
  g_nccl.Recv = (decltype(g_nccl.Recv))sym("ncclRecv");
  g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
  g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
  g_nccl_ok = ok;
}

const NcclApi& nccl() {
  std::call_once(g_nccl_once, load_nccl);
  if (!g_nccl_ok) SPB_FAIL(SPB_NCCL_ERROR, "libnccl.so.2 could not be loaded");
  return g_nccl;
}

int Ctx::world() const { return dist ? dist->world : 1; }
int Ctx::rank() const { return dist ? dist->rank : 0; }

__global__ void peer_allreduce_kernel(double* v, int count, PeerPtrs pp) {
  __shared__ double loc[8];
  if (threadIdx.x < 4) {  // plain doubles: (hi, lo) = (v, 0)
    loc[2 * threadIdx.x] = (int)threadIdx.x < count ? v[threadIdx.x] : 0.0;
    loc[2 * threadIdx.x + 1] = 0.0;
  }
  peer_allreduce_dd(loc, pp);
  if ((int)threadIdx.x < count) v[threadIdx.x] = loc[threadIdx.x];
}

void allreduce_sum(Ctx* ctx, double* dev, size_t count) {
  if (!ctx->dist || ctx->dist->world == 1) return;
  if (ctx->dist->peer) {
    if (count > 4) SPB_FAIL(SPB_INVALID_ARG, "peer all-reduce carries at most 4 doubles");
    LaunchScope ls(ctx, FAM_SCALAR);
    peer_allreduce_kernel<<<1, 32, 0, ctx->stream>>>(dev, (int)count, ctx->dist->scal->ptrs());
    check_launch("peer_allreduce_kernel");
    return;
  }
  SPB_NCCL(nccl().AllReduce(dev, dev, count, ncclFloat64, ncclSum, ctx->dist->comm, ctx->stream));
}

// ---------------------------------------------------------------- peer windows (CUDA IPC)
PeerWindow* window_create(Ctx* ctx, size_t bytes) {
  Dist* d = ctx->dist;
  const int W = d->world, me = d->rank;
  if (W > kMaxPeers) return nullptr;
  auto* w = new PeerWindow();
  w->world = W;
  w->rank = me;
  w->bytes = (bytes + 255) & ~(size_t)255;
  w->mapped.assign(W, nullptr);
  bool ok = true;
  cudaIpcMemHandle_t h;
  memset(&h, 0, sizeof(h));
  if (cudaMalloc(&w->local, w->bytes) != cudaSuccess) {
    cudaGetLastError();
    w->local = nullptr;
    ok = false;
  }
  if (ok) {
    SPB_CUDA(cudaMemsetAsync(w->local, 0, w->bytes, ctx->stream));
    SPB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (cudaIpcGetMemHandle(&h, w->local) != cudaSuccess) {
      cudaGetLastError();
      ok = false;
    }
  }
  // exchange the 64-byte handles (+ an ok word) through the set-up communicator; this is also the
  // barrier that orders every rank's memset before any peer store
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  int64_t mine[9];
  memcpy(mine, &h, 64);
  mine[8] = ok ? 1 : 0;
  std::vector<int64_t> all;
  allgather_i64(ctx, mine, 9, all);
  for (int q = 0; q < W; ++q) ok = ok && all[(size_t)q * 9 + 8] == 1;
  if (ok) {
    for (int q = 0; q < W; ++q) {
      if (q == me) {
        w->mapped[q] = w->local;
        continue;
      }
      cudaIpcMemHandle_t hq;
      memcpy(&hq, &all[(size_t)q * 9], 64);
      if (cudaIpcOpenMemHandle(&w->mapped[q], hq, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        w->mapped[q] = nullptr;
        ok = false;
      }
    }
  }
  // agree on the outcome (a rank that failed to map one peer disables the transport for everyone)
  int64_t okw = ok ? 1 : 0;
  allgather_i64(ctx, &okw, 1, all);
  for (int q = 0; q < W; ++q) ok = ok && all[q] == 1;
  if (!ok) {
    window_destroy(w);
    return nullptr;
  }
  return w;
}

void window_destroy(PeerWindow* w) {
  if (!w) return;
  for (int q = 0; q < (int)w->mapped.size(); ++q)
    if (q != w->rank && w->mapped[q]) cudaIpcCloseMemHandle(w->mapped[q]);
  if (w->local) cudaFree(w->local);
  delete w;
}

void device_check(Ctx* ctx) {
  int err = 0, dev = 0;
  if (peer_mode(ctx)) {
    ScalWin* me = static_cast<ScalWin*>(ctx->dist->scal->local);
    SPB_CUDA(cudaMemcpyAsync(&err, &me->error, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (ctx->dev_err) SPB_CUDA(cudaMemcpyAsync(&dev, ctx->dev_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SPB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (err || dev) {  // report once: clear the words so that the context stays usable for diagnostics
    if (err) cudaMemsetAsync(&static_cast<ScalWin*>(ctx->dist->scal->local)->error, 0, sizeof(int), ctx->stream);
    if (dev) cudaMemsetAsync(ctx->dev_err, 0, sizeof(int), ctx->stream);
    cudaStreamSynchronize(ctx->stream);
  }
  if (err) SPB_FAIL(SPB_NCCL_ERROR, "timed out waiting for a peer rank (scalar all-reduce over peer memory)");
  if (dev == 1) SPB_FAIL(SPB_NCCL_ERROR, "timed out waiting for a neighbour's halo put (peer memory)");
  if (dev) SPB_FAIL(SPB_CUDA_ERROR, "Gauss-Seidel wavefront sweep: timed out waiting for a value of another block");
}

void allgather_i64(Ctx* ctx, const int64_t* host_in, size_t count, std::vector<int64_t>& out) {
  const int w = ctx->world();
  out.resize(count * w);
  if (w == 1) {
    std::copy(host_in, host_in + count, out.begin());
    return;
  }
  Dist* d = ctx->dist;
  d->scratch.ensure(sizeof(int64_t) * count * (w + 1));
  int64_t* send = bufptr<int64_t>(d->scratch);
  int64_t* recv = send + count;
  SPB_CUDA(cudaMemcpyAsync(send, host_in, sizeof(int64_t) * count, cudaMemcpyHostToDevice, ctx->stream));
  SPB_NCCL(nccl().AllGather(send, recv, count, ncclInt64, d->comm, ctx->stream));
  SPB_CUDA(cudaMemcpyAsync(out.data(), recv, sizeof(int64_t) * count * w, cudaMemcpyDeviceToHost, ctx->stream));
  SPB_CUDA(cudaStreamSynchronize(ctx->stream));
}

// ---------------------------------------------------------------- localisation kernels
__global__ void mark_halo_kernel(const int* cols, int64_t nnz, int rb, int re, unsigned* bitmap) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz;
       k += (int64_t)gridDim.x * blockDim.x) {
    const int c = cols[k];
    if (c < rb || c >= re) atomicOr(&bitmap[c >> 5], 1u << (c & 31));  // integer atomic
  }
}
__global__ void popc_kernel(const unsigned* bitmap, int64_t nwords, int* counts) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < nwords) counts[i] = __popc(bitmap[i]);
  if (i == nwords) counts[i] = 0;
}
__global__ void enumerate_halo_kernel(const unsigned* bitmap, const int* offsets, int64_t nwords,
                                      int* halo_cols) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nwords) return;
  unsigned b = bitmap[i];
  int o = offsets[i];
  while (b) {
    const int bit = __ffs(b) - 1;
    halo_cols[o++] = (int)(i * 32 + bit);
    b &= b - 1;
  }
}
__global__ void remap_cols_kernel(int* cols, int64_t nnz, int rb, int re, const unsigned* bitmap,
                                  const int* offsets) {
  const int nl = re - rb;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz;
       k += (int64_t)gridDim.x * blockDim.x) {
    const int c = cols[k];
    if (c >= rb && c < re) {
      cols[k] = c - rb;
    } else {
      const int w = c >> 5, b = c & 31;
      cols[k] = nl + offsets[w] + __popc(bitmap[w] & ((1u << b) - 1u));
    }
  }
}
__global__ void shift_idx_kernel(int* idx, int64_t n, int rb) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) idx[i] -= rb;
}
template <typename IP>
__global__ void tile_boundary_kernel(const IP* indptr, const int* cols, const int* tile_row,
                                     int64_t ntiles, int n_local, unsigned char* flags) {
  // one warp per tile
  const int64_t t = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (t >= ntiles) return;
  const IP s = indptr[tile_row[t]], e = indptr[tile_row[t + 1]];
  int f = 0;
  for (IP k = s + lane; k < e; k += 32) f |= (cols[k] >= n_local);
  f = __any_sync(0xffffffffu, f);
  if (lane == 0) flags[t] = (unsigned char)f;
}
template <typename T>
__global__ void pack_kernel(const T* x, const int* idx, int64_t n, T* out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = x[idx[i]];
}

// Peer transport: gather this rank's boundary entries of x and store them straight into the
// neighbours' halo windows over NVLink; the CTA that finishes last publishes the sequence number
// to every neighbour (release, system scope).
template <typename T>
__global__ void halo_put_kernel(const T* x, const int* idx, long long total, HaloHead* head, PutArgs pa, int* err) {
  const unsigned long long seq = *((volatile unsigned long long*)&head->seq) + 1;
  const long long par = (long long)(seq & 1);
  if (threadIdx.x == 0) {
    // this rank's SpMV seq-1 is complete (stream order): acknowledge it to every peer, THEN wait until
    // every peer has consumed exchange seq-2, the previous user of the parity buffer written below
    // (publish before waiting: no cycle).  Every CTA waits; block 0 publishes.
    if (blockIdx.x == 0)
      for (int j = 0; j < pa.npeers; ++j) st_release_sys(pa.rack[j], seq - 1);
    if (seq >= 3)
      for (int j = 0; j < pa.npeers; ++j)
        if (!spin_until_ge(&head->acks[head->peer_rank[j]], seq - 2)) {
          head->error = 1;
          if (err) *err = 1;
        }
  }
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int j = 0;
    while (j + 1 < pa.npeers && i >= pa.send_off[j + 1]) ++j;
    T* dst = static_cast<T*>(pa.dst0[j]) + par * pa.dst_stride[j] + (i - pa.send_off[j]);
    *dst = x[idx[i]];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(&head->done, 1u);
    if (prev == gridDim.x - 1) {
      __threadfence_system();
      head->done = 0;
      for (int j = 0; j < pa.npeers; ++j) st_release_sys(pa.rflag[j], seq);
      *((volatile unsigned long long*)&head->seq) = seq;
    }
  }
}

template <typename T>
void halo_put(CsrMat<T>* m, const T* x) {
  Ctx* c = m->ctx;
  const long long total = m->put.send_off[m->put.npeers];
  LaunchScope ls(c, FAM_PACK);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, 256), 148 * 4));
  halo_put_kernel<T><<<grid, 256, 0, c->stream>>>(x, bufptr<int>(m->send_idx), total,
                                                   static_cast<HaloHead*>(m->halo_win->local), m->put, c->dev_err);
  check_launch("halo_put_kernel");
}

template <typename T>
void halo_release(CsrMat<T>* m) {
  if (m->halo_win) {
    cudaStreamSynchronize(m->ctx->stream);
    window_destroy(m->halo_win);
    m->halo_win = nullptr;
  }
}

template <typename T>
void csr_localize(CsrMat<T>* m) {
  Ctx* c = m->ctx;
  Dist* d = c->dist;
  const int W = d->world, me = d->rank;
  const int rb = (int)m->row_begin, re = (int)(m->row_begin + m->n_local);
  std::vector<int64_t> starts;
  {
    int64_t mine = m->row_begin;
    allgather_i64(c, &mine, 1, starts);
    starts.push_back(m->n_global);
    for (int p = 0; p < W; ++p)
      if (starts[p] > starts[p + 1]) SPB_FAIL(SPB_INVALID_ARG, "row blocks must be ordered by rank");
  }
  const int64_t nwords = ceil_div(m->n_global, 32);
  DevBuf bitmap, counts, offsets;
  bitmap.alloc(sizeof(unsigned) * nwords);
  counts.alloc(sizeof(int) * (nwords + 1));
  offsets.alloc(sizeof(int) * (nwords + 1));
  SPB_CUDA(cudaMemsetAsync(bitmap.p, 0, bitmap.bytes, c->stream));
  const int gnnz = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(m->nnz, 256), 148 * 16));
  mark_halo_kernel<<<gnnz, 256, 0, c->stream>>>(bufptr<int>(m->cols), m->nnz, rb, re, bufptr<unsigned>(bitmap));
  popc_kernel<<<(int)ceil_div(nwords + 1, 256), 256, 0, c->stream>>>(bufptr<unsigned>(bitmap), nwords, bufptr<int>(counts));
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, bufptr<int>(counts), bufptr<int>(offsets), nwords + 1, c->stream);
  DevBuf tmp;
  tmp.alloc(tb);
  cub::DeviceScan::ExclusiveSum(tmp.p, tb, bufptr<int>(counts), bufptr<int>(offsets), nwords + 1, c->stream);
  int nh = 0;
  SPB_CUDA(cudaMemcpyAsync(&nh, bufptr<int>(offsets) + nwords, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  SPB_CUDA(cudaStreamSynchronize(c->stream));
  check_launch("localize");
  m->n_halo = nh;
  DevBuf halo_cols;
  halo_cols.alloc(sizeof(int) * (size_t)std::max(nh, 1));
  enumerate_halo_kernel<<<(int)ceil_div(nwords, 256), 256, 0, c->stream>>>(
      bufptr<unsigned>(bitmap), bufptr<int>(offsets), nwords, bufptr<int>(halo_cols));
  remap_cols_kernel<<<gnnz, 256, 0, c->stream>>>(bufptr<int>(m->cols), m->nnz, rb, re, bufptr<unsigned>(bitmap), bufptr<int>(offsets));
  m->halo_cols_global.resize(nh);
  if (nh) SPB_CUDA(cudaMemcpyAsync(m->halo_cols_global.data(), halo_cols.p, sizeof(int) * nh, cudaMemcpyDeviceToHost, c->stream));
  SPB_CUDA(cudaStreamSynchronize(c->stream));
  check_launch("localize2");

  // who owns what I need: halo slots are sorted by global id, hence grouped by owner rank
  std::vector<int64_t> need(W, 0), need_off(W + 1, 0);
  {
    int p = 0;
    for (int i = 0; i < nh; ++i) {
      const int64_t g = m->halo_cols_global[i];
      while (g >= starts[p + 1]) ++p;
      need[p]++;
    }
    for (int p2 = 0; p2 < W; ++p2) need_off[p2 + 1] = need_off[p2] + need[p2];
  }
  std::vector<int64_t> all_need;
  allgather_i64(c, need.data(), W, all_need);  // all_need[p*W + q] = what p needs from q
  int64_t total_send = 0;
  m->peers.clear();
  for (int q = 0; q < W; ++q) {
    if (q == me) continue;
    const int64_t rc = need[q], sc = all_need[(size_t)q * W + me];
    if (rc == 0 && sc == 0) continue;
    HaloPeer hp{q, total_send, sc, need_off[q], rc};
    total_send += sc;
    m->peers.push_back(hp);
  }
  if (d->peer) {
    // halo window: head + 2 x n_halo T (parity); every rank learns every n_halo to address parity 1
    if ((int)m->peers.size() > kMaxPeers) SPB_FAIL(SPB_INVALID_ARG, "too many halo peers");
    int64_t nh64 = nh;
    std::vector<int64_t> nh_all;
    allgather_i64(c, &nh64, 1, nh_all);
    m->halo_win = window_create(c, kHaloHeadBytes + 2 * sizeof(T) * (size_t)std::max<int64_t>(nh, 1));
    if (!m->halo_win) SPB_FAIL(SPB_NCCL_ERROR, "could not map the halo window of a peer (CUDA IPC); set SPB_COMM=nccl");
    HaloHead hh;
    memset(&hh, 0, sizeof(hh));
    PutArgs& pa = m->put;
    memset(&pa, 0, sizeof(pa));
    pa.npeers = (int)m->peers.size();
    hh.npeers = pa.npeers;
    for (int j = 0; j < pa.npeers; ++j) {
      const HaloPeer& hp = m->peers[j];
      const int q = hp.rank;
      hh.peer_rank[j] = q;
      int64_t roff = 0;  // where my entries start in q's halo: q's slots are grouped by owner rank
      for (int r = 0; r < me; ++r) roff += all_need[(size_t)q * W + r];
      char* qbase = static_cast<char*>(m->halo_win->mapped[q]);
      pa.dst0[j] = qbase + kHaloHeadBytes + sizeof(T) * (size_t)roff;
      pa.dst_stride[j] = std::max<int64_t>(nh_all[q], 1);
      pa.rflag[j] = &reinterpret_cast<HaloHead*>(qbase)->flags[me];
      pa.rack[j] = &reinterpret_cast<HaloHead*>(qbase)->acks[me];
      pa.send_off[j] = hp.send_off;
    }
    pa.send_off[pa.npeers] = total_send;
    SPB_CUDA(cudaMemcpyAsync(m->halo_win->local, &hh, sizeof(hh), cudaMemcpyHostToDevice, c->stream));
    SPB_CUDA(cudaStreamSynchronize(c->stream));
    int64_t token = 1;  // nobody starts putting before every head is initialised
    std::vector<int64_t> sink;
    allgather_i64(c, &token, 1, sink);
  } else {
    m->halo.alloc(sizeof(T) * (size_t)std::max<int64_t>(nh, 1));
  }
  m->sendbuf.alloc(sizeof(T) * (size_t)std::max<int64_t>(total_send, 1));
  m->send_idx.alloc(sizeof(int) * (size_t)std::max<int64_t>(total_send, 1));
  // exchange the request lists (global ids), then make them local
  SPB_NCCL(nccl().GroupStart());
  for (const HaloPeer& hp : m->peers) {
    if (hp.recv_cnt)
      SPB_NCCL(nccl().Send(bufptr<int>(halo_cols) + hp.recv_off, hp.recv_cnt, ncclInt32, hp.rank, d->comm_halo, c->stream));
    if (hp.send_cnt)
      SPB_NCCL(nccl().Recv(bufptr<int>(m->send_idx) + hp.send_off, hp.send_cnt, ncclInt32, hp.rank, d->comm_halo, c->stream));
  }
  SPB_NCCL(nccl().GroupEnd());
  if (total_send)
    shift_idx_kernel<<<(int)ceil_div(total_send, 256), 256, 0, c->stream>>>(bufptr<int>(m->send_idx), total_send, rb);
  SPB_CUDA(cudaStreamSynchronize(c->stream));
  check_launch("localize3");
}

template <typename T>
void classify_tiles(CsrMat<T>* m) {
  Ctx* c = m->ctx;
  DevBuf flags;
  flags.alloc((size_t)m->ntiles);
  const int grid = (int)ceil_div(m->ntiles * 32, 256);
  if (m->ip64)
    tile_boundary_kernel<int64_t><<<grid, 256, 0, c->stream>>>(bufptr<int64_t>(m->indptr), bufptr<int>(m->cols), bufptr<int>(m->tile_row), m->ntiles, (int)m->n_local, bufptr<unsigned char>(flags));
  else
    tile_boundary_kernel<int32_t><<<grid, 256, 0, c->stream>>>(bufptr<int32_t>(m->indptr), bufptr<int>(m->cols), bufptr<int>(m->tile_row), m->ntiles, (int)m->n_local, bufptr<unsigned char>(flags));
  check_launch("tile_boundary_kernel");
  std::vector<unsigned char> h((size_t)m->ntiles);
  SPB_CUDA(cudaMemcpyAsync(h.data(), flags.p, h.size(), cudaMemcpyDeviceToHost, c->stream));
  SPB_CUDA(cudaStreamSynchronize(c->stream));
  std::vector<int> ti, tb;
  for (int64_t t = 0; t < m->ntiles; ++t) (h[t] ? tb : ti).push_back((int)t);
  m->n_tiles_interior = (int64_t)ti.size();
  m->n_tiles_boundary = (int64_t)tb.size();
  m->tiles_interior.alloc(sizeof(int) * std::max<size_t>(ti.size(), 1));
  m->tiles_boundary.alloc(sizeof(int) * std::max<size_t>(tb.size(), 1));
  if (!ti.empty()) SPB_CUDA(cudaMemcpyAsync(m->tiles_interior.p, ti.data(), sizeof(int) * ti.size(), cudaMemcpyHostToDevice, c->stream));
  if (!tb.empty()) SPB_CUDA(cudaMemcpyAsync(m->tiles_boundary.p, tb.data(), sizeof(int) * tb.size(), cudaMemcpyHostToDevice, c->stream));
  std::vector<int> all(ti);
  all.insert(all.end(), tb.begin(), tb.end());
  m->tiles_all.alloc(sizeof(int) * std::max<size_t>(all.size(), 1));
  if (!all.empty()) SPB_CUDA(cudaMemcpyAsync(m->tiles_all.p, all.data(), sizeof(int) * all.size(), cudaMemcpyHostToDevice, c->stream));
  SPB_CUDA(cudaStreamSynchronize(c->stream));
}

template <typename T>
void halo_exchange_begin(CsrMat<T>* m, const T* x) {
  Ctx* c = m->ctx;
  Dist* d = c->dist;
  int64_t total_send = 0;
  for (const HaloPeer& hp : m->peers) total_send += hp.send_cnt;
  if (total_send) {
    LaunchScope ls(c, FAM_PACK);
    const int grid = (int)std::min<int64_t>(ceil_div(total_send, 256), 148 * 8);
    pack_kernel<T><<<grid, 256, 0, c->stream>>>(x, bufptr<int>(m->send_idx), total_send, bufptr<T>(m->sendbuf));
    check_launch("pack_kernel");
  }
  SPB_CUDA(cudaEventRecord(c->ev_pack, c->stream));
  SPB_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_pack, 0));
  // payload in units of the real type: f64 / f32 words
  const size_t per = sizeof(T) / sizeof(real_t<T>);
  const ncclDataType_t ndt = sizeof(real_t<T>) == 4 ? ncclFloat32 : ncclFloat64;
  SPB_NCCL(nccl().GroupStart());
  for (const HaloPeer& hp : m->peers) {
    if (hp.send_cnt)
      SPB_NCCL(nccl().Send(bufptr<T>(m->sendbuf) + hp.send_off, hp.send_cnt * per, ndt, hp.rank, d->comm_halo, c->comm_stream));
    if (hp.recv_cnt)
      SPB_NCCL(nccl().Recv(bufptr<T>(m->halo) + hp.recv_off, hp.recv_cnt * per, ndt, hp.rank, d->comm_halo, c->comm_stream));
  }
  SPB_NCCL(nccl().GroupEnd());
  SPB_CUDA(cudaEventRecord(c->ev_halo, c->comm_stream));
}

template <typename T>
void halo_exchange_wait(CsrMat<T>* m) {
  SPB_CUDA(cudaStreamWaitEvent(m->ctx->stream, m->ctx->ev_halo, 0));
}

#define SPB_INST_DIST(T)                                              \
  template void csr_localize<T>(CsrMat<T>*);                          \
  template void classify_tiles<T>(CsrMat<T>*);                        \
  template void halo_exchange_begin<T>(CsrMat<T>*, const T*);         \
  template void halo_exchange_wait<T>(CsrMat<T>*);                    \
  template void halo_put<T>(CsrMat<T>*, const T*);                    \
  template void halo_release<T>(CsrMat<T>*);
SPB_INST_DIST(double)
SPB_INST_DIST(cplx)
SPB_INST_DIST(float)
SPB_INST_DIST(cplxf)

}  // namespace spb
