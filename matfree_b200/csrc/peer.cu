// Multi-GPU plumbing over peer memory (one process per GPU, NVLink / NVSwitch):
//   * mf_comm_*: every rank allocates one region [PeerControl | heap] with cudaMalloc, exports it
//     as a cudaIpcMemHandle_t (the host layer all-gathers the 64-byte handles through whatever
//     channel it has: torch.distributed, MPI, a file) and maps the regions of its peers;
//   * halo_push_kernel: the boundary rows of a row-sharded Lanczos block are stored straight into
//     the neighbours' extended blocks, followed by a release/acquire flag handshake, so the
//     SpMM that follows on the stream finds its halo filled -- no NCCL send/recv, no host;
//   * peer_barrier_kernel: flag barrier over the communicator (used where a halo buffer is reused
//     without an intervening reduction).
// The all-reduce itself is not here: it is fused into the tail of every reducing kernel
// (common.cuh, finalize_if_last).
//
// Replaces what XLA would insert for the reference under a row-sharded `jax.Array`: the
// collective-permute of the halo before the matvec (matfree/decomp.py:460, :287) and the
// all-reduce after every `linalg.inner` / `vector_norm` (decomp.py:288,290,463,468,471).
#include <mutex>
#include <new>

#include "internal.h"

struct mf_comm {
  int world = 0, rank = 0;
  int64_t heap_bytes = 0;
  unsigned char* region = nullptr;                 // local [PeerControl | heap]
  unsigned char* mapped[mf::kMaxPeers] = {};       // region of every rank in this address space
  mf::PeerCtx* dev_ctx = nullptr;                  // device copy of the descriptor
  bool connected = false;
};

namespace mf {
namespace {

constexpr int64_t kHeapOffset = (int64_t)((sizeof(PeerControl) + 255) / 256 * 256);

struct HaloSegs {
  int nsend;
  int peer[kMaxHaloSends];
  int64_t src_off[kMaxHaloSends];  // bytes from the local extended block
  int64_t dst_off[kMaxHaloSends];  // bytes from the peer's extended block
  int64_t bytes[kMaxHaloSends];
  int nrecv;
  int recv_peer[kMaxPeers];
};

// All CTAs copy (grid-stride over the concatenated segments); the CTA that draws the last
// ticket publishes: fence, one flag per destination, then waits for the flags of the ranks this
// rank receives from.  When the kernel completes, this rank's halo rows are in place.
__global__ void __launch_bounds__(256)
halo_push_kernel(const PeerCtx* __restrict__ pc, int64_t ext_off, HaloSegs segs) {
  const int rank = pc->rank;
  const unsigned char* src_base = pc->heap[rank] + ext_off;
  for (int s = 0; s < segs.nsend; ++s) {
    const unsigned char* src = src_base + segs.src_off[s];
    unsigned char* dst = pc->heap[segs.peer[s]] + ext_off + segs.dst_off[s];
    const int64_t bytes = segs.bytes[s];
    if ((((uintptr_t)src | (uintptr_t)dst | (uintptr_t)bytes) & 15) == 0) {
      const int64_t nv = bytes >> 4;
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
           i += (int64_t)gridDim.x * blockDim.x)
        reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
    } else {  // element size is at least 4 bytes
      const int64_t nv = bytes >> 2;
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
           i += (int64_t)gridDim.x * blockDim.x)
        reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
    }
  }
  PeerControl* mine = pc->ctl[rank];
  __shared__ bool is_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(&mine->halo_ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence_system();
  const unsigned int seq = *(volatile unsigned int*)&mine->halo_seq + 1u;
  if (threadIdx.x < segs.nsend) {
    // several segments may go to the same peer: the flag value is the same, so this is idempotent
    st_release_sys(&pc->ctl[segs.peer[threadIdx.x]]->halo_flag[rank], seq);
  }
  if (threadIdx.x < segs.nrecv)
    wait_flag(&mine->halo_flag[segs.recv_peer[threadIdx.x]], seq, &mine->error);
  __syncthreads();
  if (threadIdx.x == 0) {
    mine->halo_seq = seq;
    mine->halo_ticket = 0u;
  }
}

__global__ void peer_barrier_kernel(const PeerCtx* __restrict__ pc) {
  const int rank = pc->rank, world = pc->world;
  PeerControl* mine = pc->ctl[rank];
  const unsigned int seq = *(volatile unsigned int*)&mine->bar_seq + 1u;
  __threadfence_system();
  if ((int)threadIdx.x < world) {
    st_release_sys(&pc->ctl[threadIdx.x]->bar_flag[rank], seq);
    wait_flag(&mine->bar_flag[threadIdx.x], seq, &mine->error);
  }
  __syncthreads();
  if (threadIdx.x == 0) mine->bar_seq = seq;
}

int32_t cuda_fail(const char* what, cudaError_t e) {
  cudaGetLastError();
  set_error("%s: %s", what, cudaGetErrorString(e));
  return MF_ERR_CUDA;
}

}  // namespace

const PeerCtx* comm_ctx(const mf_comm* c) {
  return (c != nullptr && c->connected && c->world > 1) ? c->dev_ctx : nullptr;
}
void* comm_heap(const mf_comm* c) { return c ? c->region + kHeapOffset : nullptr; }

int32_t launch_halo_exchange(const mf_comm* c, const mf_halo_plan_t* plan, int64_t heap_offset,
                             int64_t block_index, int64_t ld, int32_t dtype, cudaStream_t st) {
  if (c == nullptr || !c->connected || c->world <= 1 || plan == nullptr) return MF_OK;
  if (plan->num_sends == 0 && plan->num_recv_peers == 0) return MF_OK;
  if (plan->num_sends > kMaxHaloSends || plan->num_recv_peers > kMaxPeers || plan->num_sends < 0 ||
      plan->num_recv_peers < 0) {
    set_error("halo_exchange: at most %d send segments and %d source ranks", kMaxHaloSends,
              kMaxPeers);
    return MF_ERR_INVALID_ARGUMENT;
  }
  MF_KSCOPE(MF_KC_OTHER, st);
  const int64_t rowb = ld * (int64_t)dtype_size(dtype);
  HaloSegs segs{};
  int64_t total = 0;
  segs.nsend = plan->num_sends;
  for (int i = 0; i < plan->num_sends; ++i) {
    const mf_halo_send_t& s = plan->sends[i];
    if (s.peer < 0 || s.peer >= c->world || s.peer == c->rank || s.rows < 0 || s.src_row < 0 ||
        s.dst_row < 0 || s.src_row + s.rows > plan->rows_alloc ||
        s.dst_row + s.rows > s.dst_rows_alloc ||
        heap_offset + (block_index + 1) * plan->rows_alloc * rowb > c->heap_bytes ||
        heap_offset + (block_index + 1) * s.dst_rows_alloc * rowb > c->heap_bytes) {
      set_error("halo_exchange: bad send segment %d", i);
      return MF_ERR_INVALID_ARGUMENT;
    }
    segs.peer[i] = s.peer;
    segs.src_off[i] = (block_index * plan->rows_alloc + s.src_row) * rowb;
    segs.dst_off[i] = (block_index * s.dst_rows_alloc + s.dst_row) * rowb;
    segs.bytes[i] = s.rows * rowb;
    total += segs.bytes[i];
  }
  segs.nrecv = plan->num_recv_peers;
  for (int i = 0; i < plan->num_recv_peers; ++i) segs.recv_peer[i] = plan->recv_peers[i];
  int64_t want = (total / 16 + 255) / 256;
  int grid = (int)(want < 1 ? 1 : (want > 2 * num_sms() ? 2 * num_sms() : want));
  halo_push_kernel<<<grid, 256, 0, st>>>(c->dev_ctx, heap_offset, segs);
  return check_launch("halo_push");
}

int32_t launch_peer_barrier(const mf_comm* c, cudaStream_t st) {
  if (c == nullptr || !c->connected || c->world <= 1) return MF_OK;
  MF_KSCOPE(MF_KC_OTHER, st);
  peer_barrier_kernel<<<1, 32, 0, st>>>(c->dev_ctx);
  return check_launch("peer_barrier");
}

}  // namespace mf

using namespace mf;

extern "C" {

int32_t mf_comm_create(int32_t world, int32_t rank, int64_t heap_bytes, mf_comm_t** out) {
  if (out == nullptr || world < 1 || world > kMaxPeers || rank < 0 || rank >= world ||
      heap_bytes < 0) {
    set_error("comm_create: world must be in [1, %d], rank in [0, world), heap_bytes >= 0",
              kMaxPeers);
    return MF_ERR_INVALID_ARGUMENT;
  }
  mf_comm* c = new (std::nothrow) mf_comm();
  if (c == nullptr) return MF_ERR_CUDA;
  c->world = world;
  c->rank = rank;
  c->heap_bytes = (heap_bytes + 255) / 256 * 256;
  cudaError_t e = cudaMalloc((void**)&c->region, (size_t)(kHeapOffset + c->heap_bytes));
  if (e != cudaSuccess) {
    delete c;
    return cuda_fail("comm_create: cudaMalloc of the peer region", e);
  }
  // control block zeroed (flags, sequence numbers); the heap is left to its users
  e = cudaMemset(c->region, 0, (size_t)kHeapOffset);
  if (e == cudaSuccess) e = cudaMalloc((void**)&c->dev_ctx, sizeof(PeerCtx));
  if (e != cudaSuccess) {
    cudaFree(c->region);
    delete c;
    return cuda_fail("comm_create", e);
  }
  c->mapped[rank] = c->region;
  *out = c;
  return MF_OK;
}

int32_t mf_comm_handle(const mf_comm_t* c, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == MF_COMM_HANDLE_BYTES, "IPC handle size");
  if (c == nullptr || handle64 == nullptr) {
    set_error("comm_handle: bad arguments");
    return MF_ERR_INVALID_ARGUMENT;
  }
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, c->region);
  if (e != cudaSuccess) return cuda_fail("comm_handle: cudaIpcGetMemHandle", e);
  memcpy(handle64, &h, sizeof(h));
  return MF_OK;
}

int32_t mf_comm_connect(mf_comm_t* c, const void* handles) {
  if (c == nullptr || (handles == nullptr && c->world > 1)) {
    set_error("comm_connect: bad arguments");
    return MF_ERR_INVALID_ARGUMENT;
  }
  if (c->connected) return MF_OK;
  for (int p = 0; p < c->world; ++p) {
    if (p == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)p * MF_COMM_HANDLE_BYTES, sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return cuda_fail("comm_connect: cudaIpcOpenMemHandle", e);
    c->mapped[p] = (unsigned char*)ptr;
  }
  PeerCtx ctx{};
  ctx.world = c->world;
  ctx.rank = c->rank;
  for (int p = 0; p < c->world; ++p) {
    ctx.ctl[p] = reinterpret_cast<PeerControl*>(c->mapped[p]);
    ctx.heap[p] = c->mapped[p] + kHeapOffset;
  }
  cudaError_t e = cudaMemcpy(c->dev_ctx, &ctx, sizeof(ctx), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return cuda_fail("comm_connect: descriptor upload", e);
  c->connected = true;
  return MF_OK;
}

void* mf_comm_heap(const mf_comm_t* c) { return comm_heap(c); }
int64_t mf_comm_heap_bytes(const mf_comm_t* c) { return c ? c->heap_bytes : 0; }

int32_t mf_comm_status(const mf_comm_t* c, void* stream) {
  if (c == nullptr) return MF_ERR_INVALID_ARGUMENT;
  unsigned int err = 0;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemcpyAsync(&err, c->region + offsetof(PeerControl, error), sizeof(err),
                                  cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail("comm_status", e);
  if (err != 0) {
    set_error("a peer-memory wait timed out: a rank of the communicator did not arrive");
    return MF_ERR_PEER_TIMEOUT;
  }
  return MF_OK;
}

int32_t mf_comm_disconnect(mf_comm_t* c) {
  if (c == nullptr) return MF_OK;
  for (int p = 0; p < c->world; ++p)
    if (p != c->rank && c->mapped[p] != nullptr) {
      cudaIpcCloseMemHandle(c->mapped[p]);
      c->mapped[p] = nullptr;
    }
  c->connected = false;
  cudaGetLastError();
  return MF_OK;
}

int32_t mf_comm_destroy(mf_comm_t* c) {
  if (c == nullptr) return MF_OK;
  mf_comm_disconnect(c);
  if (c->dev_ctx) cudaFree(c->dev_ctx);
  if (c->region) cudaFree(c->region);
  cudaGetLastError();
  delete c;
  return MF_OK;
}

int32_t mf_comm_barrier(const mf_comm_t* c, void* stream) {
  return launch_peer_barrier(c, (cudaStream_t)stream);
}

int32_t mf_halo_exchange(const mf_comm_t* c, const mf_halo_plan_t* plan, int64_t heap_offset,
                         int64_t block_index, int64_t ld, int32_t dtype, int32_t barrier_first,
                         void* stream) {
  if (c == nullptr || plan == nullptr || !valid_ld(ld) || heap_offset < 0 || heap_offset % 16 ||
      block_index < 0 ||
      (dtype != MF_F32 && dtype != MF_F64)) {
    set_error("halo_exchange: bad arguments");
    return MF_ERR_INVALID_ARGUMENT;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (barrier_first) MF_TRY(launch_peer_barrier(c, st));
  return launch_halo_exchange(c, plan, heap_offset, block_index, ld, dtype, st);
}

}  // extern "C"
