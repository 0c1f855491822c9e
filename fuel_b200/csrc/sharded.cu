// sharded.cu -- the z-sharded ESDF update across GPUs (BASELINE config 4) behind the C ABI.
//
// Multi-GPU form of SDFMap::updateESDF3d (plan_env/src/sdf_map.cpp:152-241) over the whole map: rank r owns the
// planes [r*nz/G, (r+1)*nz/G) of every (x,y) column, z fastest like the reference (sdf_map.h:145-147).  The
// squared EDT is separable and exact in integers, so the sweeps may run wherever their lines are whole:
//   1. all-to-all of the occupancy byte (1 B/voxel): z-slabs -> x-slabs, so that rank r sees whole z lines of its
//      x range (G chunks of nz/G planes each);
//   2. zpack + zy tiles on the x-slab (esdf_tile.cu), written straight into the send layout of step 3; the tiles
//      are produced destination by destination;
//   3. THE exchange of the 2-D partial (int32): x-slabs -> z-slabs.  Peer-memory form (default where the ranks can
//      map each other's memory: one process per GPU, NVLink/NVSwitch): there is no exchange step at all -- the zy
//      tile kernel of destination d stores its rows STRAIGHT INTO rank d's receive buffer over NVLink (the buffers
//      are CUDA-IPC mapped once at creation, the handles travel through an ncclAllGather), so the transfer IS the
//      kernel's output traffic, tile by tile; one 4-byte-per-rank all-gather afterwards is the barrier that tells
//      every rank that all its sources have finished writing.  NCCL form (fallback, FUELGPU_SHARDED_P2P=0): one
//      grouped ncclSend/ncclRecv pair per round, round k (peer r+k / r-k) enqueued on a second stream as soon as
//      the tiles of destination r+k are done, so the transfer of round k overlaps the tiles of round k+1;
//   4. x tiles on the own z-slab: rows of a tile are G contiguous pieces (one per source rank), result in metres
//      in the caller's z-slab of distance_buffer_.
// NCCL is bound at run time (dlopen of libnccl.so.2, the copy already loaded by the host process if any), so
// libfuelgpu has no link-time dependency on it and single-GPU users never touch it.
#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>

#include <vector>

namespace {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
  bool ok;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.ok) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fuel_fail(nullptr, FUELGPU_EUNSUPPORTED, "libnccl.so.2 not found: %s", dlerror());
#define LOADSYM(field, name)                                                             \
  *(void**)(&g_nccl.field) = dlsym(h, name);                                             \
  if (!g_nccl.field) return fuel_fail(nullptr, FUELGPU_EUNSUPPORTED, "NCCL symbol %s missing", name)
  LOADSYM(GetUniqueId, "ncclGetUniqueId");
  LOADSYM(CommInitRank, "ncclCommInitRank");
  LOADSYM(CommDestroy, "ncclCommDestroy");
  LOADSYM(Send, "ncclSend");
  LOADSYM(Recv, "ncclRecv");
  LOADSYM(GroupStart, "ncclGroupStart");
  LOADSYM(GroupEnd, "ncclGroupEnd");
  LOADSYM(AllGather, "ncclAllGather");
  LOADSYM(GetErrorString, "ncclGetErrorString");
#undef LOADSYM
  g_nccl.ok = true;
  return 0;
}

#define FUEL_NCCL(expr)                                                                                   \
  do {                                                                                                    \
    ncclResult_t _r = (expr);                                                                             \
    if (_r != ncclSuccess) return fuel_fail(nullptr, FUELGPU_ECUDA, "NCCL: %s", g_nccl.GetErrorString(_r)); \
  } while (0)

}  // namespace

struct FuelComm {
  ncclComm_t comm;
  int nranks, rank, dev;
};

struct FuelShardedEsdf {
  FuelComm* c;
  int nx, ny, nz, nzl, nxl, NW, wl;
  double res;
  uint8_t* occ_x;  // [G][nxl][ny][nzl]: the occupancy of my x range, one chunk per source rank
  void* rec;       // z records of the x-slab
  int32_t* psend;  // [G dest][wl][ny][nxl][32]
  int32_t* precv;  // [G src][wl][ny][nxl][32]
  size_t blk;      // int32 per (rank, rank) block of the partial
  // peer-memory exchange: precv of every rank mapped here (peer_recv[r] == precv), and the barrier's scratch
  bool p2p;
  int32_t* peer_recv[64];
  int32_t* sync_buf;  // [1 + G]
  cudaStream_t comm_stream;
  cudaEvent_t ev_round[64];
  cudaEvent_t ev_comm_done;
  cudaEvent_t ev_t[5];  // stage stamps on the caller's stream
  bool timed;
};

extern "C" {

int fuelgpu_comm_get_unique_id(uint8_t id[128]) {
  if (!id) return fuel_fail(nullptr, FUELGPU_EINVAL, "null argument");
  int rc = load_nccl();
  if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId u;
  FUEL_NCCL(g_nccl.GetUniqueId(&u));
  memcpy(id, &u, 128);
  return 0;
}

int fuelgpu_comm_init(int32_t nranks, int32_t rank, const uint8_t id[128], int32_t device_id, FuelComm** out) {
  if (!id || !out || nranks < 1 || rank < 0 || rank >= nranks) return fuel_fail(nullptr, FUELGPU_EINVAL, "bad argument");
  *out = nullptr;
  int rc = load_nccl();
  if (rc) return rc;
  FUEL_CUDA(nullptr, cudaSetDevice(device_id));
  ncclUniqueId u;
  memcpy(&u, id, 128);
  FuelComm* c = new FuelComm();
  c->nranks = nranks;
  c->rank = rank;
  c->dev = device_id;
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, u, rank);
  if (r != ncclSuccess) {
    delete c;
    return fuel_fail(nullptr, FUELGPU_ECUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
  }
  *out = c;
  return 0;
}

int fuelgpu_comm_destroy(FuelComm* c) {
  if (!c) return 0;
  if (g_nccl.ok && c->comm) g_nccl.CommDestroy(c->comm);
  delete c;
  return 0;
}

int fuelgpu_comm_info(const FuelComm* c, int32_t* nranks, int32_t* rank) {
  if (!c) return fuel_fail(nullptr, FUELGPU_EINVAL, "null comm");
  if (nranks) *nranks = c->nranks;
  if (rank) *rank = c->rank;
  return 0;
}

int fuelgpu_sharded_esdf_destroy(FuelShardedEsdf* s) {
  if (!s) return 0;
  cudaSetDevice(s->c->dev);
  if (s->comm_stream) cudaStreamSynchronize(s->comm_stream);
  if (s->p2p) {
    // (collective in this mode, like the creation: nobody frees its receive buffer while a peer still maps it)
    cudaDeviceSynchronize();
    for (int g = 0; g < s->c->nranks; ++g)
      if (g != s->c->rank && s->peer_recv[g]) cudaIpcCloseMemHandle(s->peer_recv[g]);
    if (g_nccl.ok && s->c->comm && s->comm_stream &&
        g_nccl.AllGather(s->sync_buf, s->sync_buf + 1, 1, ncclInt32, s->c->comm, s->comm_stream) == ncclSuccess)
      cudaStreamSynchronize(s->comm_stream);
  }
  if (s->comm_stream) cudaStreamDestroy(s->comm_stream);
  void* ptrs[] = { s->occ_x, s->rec, s->psend, s->precv, s->sync_buf };
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (int i = 0; i < 64; ++i)
    if (s->ev_round[i]) cudaEventDestroy(s->ev_round[i]);
  if (s->ev_comm_done) cudaEventDestroy(s->ev_comm_done);
  for (int i = 0; i < 5; ++i)
    if (s->ev_t[i]) cudaEventDestroy(s->ev_t[i]);
  delete s;
  return 0;
}

int fuelgpu_sharded_esdf_create(FuelComm* c, const int32_t n[3], double resolution, FuelShardedEsdf** out) {
  if (!c || !n || !out) return fuel_fail(nullptr, FUELGPU_EINVAL, "null argument");
  *out = nullptr;
  const int G = c->nranks;
  if (G > 64) return fuel_fail(nullptr, FUELGPU_EUNSUPPORTED, "at most 64 ranks");
  for (int i = 0; i < 3; ++i)
    if (n[i] < 1 || n[i] > 1024) return fuel_fail(nullptr, FUELGPU_EINVAL, "grid extent must be in 1..1024 per axis");
  // a rank's z-slab is a whole number of 32-voxel words, its x range a whole number of 32-sample bands
  if (n[2] % (32 * G) || n[0] % (32 * G))
    return fuel_fail(nullptr, FUELGPU_EINVAL, "nx and nz must be multiples of 32 * ranks (%s%lld ranks)", "", G);
  FUEL_CUDA(nullptr, cudaSetDevice(c->dev));
  FuelShardedEsdf* s = new FuelShardedEsdf();
  memset(s, 0, sizeof(*s));
  s->c = c;
  s->nx = n[0];
  s->ny = n[1];
  s->nz = n[2];
  s->nzl = n[2] / G;
  s->nxl = n[0] / G;
  s->NW = n[2] / 32;
  s->wl = s->nzl / 32;
  s->res = resolution;
  s->blk = (size_t)s->wl * s->ny * s->nxl * 32;
  const size_t slab = (size_t)s->nxl * s->ny * s->nz;  // voxels of my x range == voxels of my z-slab
  const int NYP = (s->ny + 1) & ~1;
#define CR(expr)                                                                          \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      snprintf(g_fuelgpu_err, 512, "sharded_esdf_create: %s: %s", #expr, cudaGetErrorString(_e)); \
      fuelgpu_sharded_esdf_destroy(s);                                                    \
      return _e == cudaErrorMemoryAllocation ? FUELGPU_ENOMEM : FUELGPU_ECUDA;            \
    }                                                                                     \
  } while (0)
  CR(cudaMalloc(&s->occ_x, slab));
  CR(cudaMalloc(&s->rec, (size_t)s->nxl * s->NW * NYP * 8 + 64));
  CR(cudaMalloc(&s->psend, s->blk * G * 4));
  CR(cudaMalloc(&s->precv, s->blk * G * 4));
  CR(cudaStreamCreateWithFlags(&s->comm_stream, cudaStreamNonBlocking));
  for (int i = 0; i < G; ++i) CR(cudaEventCreateWithFlags(&s->ev_round[i], cudaEventDisableTiming));
  CR(cudaEventCreateWithFlags(&s->ev_comm_done, cudaEventDisableTiming));
  for (int i = 0; i < 5; ++i) CR(cudaEventCreate(&s->ev_t[i]));
  CR(cudaMalloc(&s->sync_buf, sizeof(int32_t) * (1 + G) + 64 * (size_t)(1 + G)));
  CR(cudaMemset(s->sync_buf, 0, sizeof(int32_t) * (1 + G)));
#undef CR
  // peer-memory exchange: every rank publishes the IPC handle of its receive buffer (all-gather over the
  // communicator), maps the others' and agrees (all ranks or none) that it worked
  s->p2p = false;
  const char* env = getenv("FUELGPU_SHARDED_P2P");
  const bool want = G > 1 && !(env && atoi(env) == 0);
  if (G > 1) {
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    int ok = want && cudaIpcGetMemHandle(&mine, s->precv) == cudaSuccess ? 1 : 0;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    char* hbuf = (char*)(s->sync_buf + 1 + G);  // [1 + G] x 64 bytes
    std::vector<cudaIpcMemHandle_t> all(G);
    cudaError_t e = cudaMemcpy(hbuf, &mine, 64, cudaMemcpyHostToDevice);
    ncclResult_t nr = ncclSuccess;
    if (e == cudaSuccess) nr = g_nccl.AllGather(hbuf, hbuf + 64, 64, ncclUint8, c->comm, s->comm_stream);
    if (e == cudaSuccess && nr == ncclSuccess) e = cudaStreamSynchronize(s->comm_stream);
    if (e == cudaSuccess && nr == ncclSuccess) e = cudaMemcpy(all.data(), hbuf + 64, 64 * (size_t)G, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess || nr != ncclSuccess) ok = 0;
    if (ok) {
      for (int g = 0; g < G && ok; ++g) {
        if (g == c->rank) {
          s->peer_recv[g] = s->precv;
          continue;
        }
        void* ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, all[g], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          ok = 0;
        } else {
          s->peer_recv[g] = (int32_t*)ptr;
        }
      }
    }
    // all or none: the minimum of the ranks' flags (gathered like the handles)
    int32_t flag = ok, flags[64];
    e = cudaMemcpy(s->sync_buf, &flag, 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) nr = g_nccl.AllGather(s->sync_buf, s->sync_buf + 1, 1, ncclInt32, c->comm, s->comm_stream);
    if (e == cudaSuccess && nr == ncclSuccess) e = cudaStreamSynchronize(s->comm_stream);
    if (e == cudaSuccess && nr == ncclSuccess) e = cudaMemcpy(flags, s->sync_buf + 1, 4 * (size_t)G, cudaMemcpyDeviceToHost);
    bool all_ok = e == cudaSuccess && nr == ncclSuccess;
    for (int g = 0; g < G && all_ok; ++g) all_ok = flags[g] == 1;
    if (all_ok) {
      s->p2p = true;
    } else {
      for (int g = 0; g < G; ++g) {
        if (g != c->rank && s->peer_recv[g]) cudaIpcCloseMemHandle(s->peer_recv[g]);
        s->peer_recv[g] = nullptr;
      }
      cudaGetLastError();
    }
  }
  *out = s;
  return 0;
}

// 1 when the partial travels by direct stores into the peers' receive buffers, 0 when it goes through ncclSend/Recv
int fuelgpu_sharded_esdf_uses_peer_memory(const FuelShardedEsdf* s) { return s && s->p2p ? 1 : 0; }

int fuelgpu_sharded_esdf_update(FuelShardedEsdf* s, void* cuda_stream, const void* occ_slab_dev, int flags,
                                void* dist_slab_dev) {
  if (!s || !occ_slab_dev || !dist_slab_dev) return fuel_fail(nullptr, FUELGPU_EINVAL, "null argument");
  if (flags & FUELGPU_ESDF_SIGNED)
    return fuel_fail(nullptr, FUELGPU_EUNSUPPORTED, "the sharded update computes the unsigned field only");
  FuelComm* c = s->c;
  const int G = c->nranks, r = c->rank;
  FUEL_CUDA(nullptr, cudaSetDevice(c->dev));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int mode = (flags & FUELGPU_ESDF_OPTIMISTIC) ? 0 : 1;
  const size_t chunk = (size_t)s->nxl * s->ny * s->nzl;  // bytes of one (x range, z-slab) block of the occupancy
  FUEL_CUDA(nullptr, cudaEventRecord(s->ev_t[0], st));
  // 1. occupancy z-slabs -> x-slabs (x is the slowest axis: the block for rank d is contiguous)
  {
    const uint8_t* src = (const uint8_t*)occ_slab_dev;
    FUEL_NCCL(g_nccl.GroupStart());
    for (int k = 0; k < G; ++k) {
      FUEL_NCCL(g_nccl.Send(src + (size_t)k * chunk, chunk, ncclUint8, k, c->comm, st));
      FUEL_NCCL(g_nccl.Recv(s->occ_x + (size_t)k * chunk, chunk, ncclUint8, k, c->comm, st));
    }
    FUEL_NCCL(g_nccl.GroupEnd());
  }
  FUEL_CUDA(nullptr, cudaEventRecord(s->ev_t[1], st));
  // 2. records of my x range (z lines assembled from the G chunks)
  int rc = edt_stage_zpack(st, s->occ_x, s->rec, s->nxl, s->ny, s->nzl, G, (int64_t)chunk, mode);
  if (rc) return fuel_fail(nullptr, rc, "zpack stage failed");
  // 2+3 (peer-memory form). ONE launch over all words: the tiles of destination d's words store into rank d's
  // receive slot of source r (my own words into my own buffer)
  if (s->p2p && G <= 16) {
    int32_t* tab[16];
    for (int d = 0; d < G; ++d) tab[d] = (d == r ? s->precv : s->peer_recv[d]) + (size_t)r * s->blk;
    rc = edt_stage_zy_scatter(st, s->rec, s->nxl, s->ny, s->NW, tab, G, s->wl, 32, (int64_t)s->ny * s->nxl * 32,
                              (int64_t)s->nxl * 32);
    if (rc) return fuel_fail(nullptr, rc, "zy stage failed");
  }
  // 2+3. zy tiles destination by destination, the peers first and this rank's own block last (it needs no
  // transfer and is written straight into the receive buffer); round k's transfer overlaps round k+1's tiles
  for (int kk = 0; kk < G && !(s->p2p && G <= 16); ++kk) {
    const int k = (kk + 1) % G;  // 1, 2, ..., G-1, 0
    const int d = (r + k) % G, src = (r - k + G) % G;
    // peer-memory form: rank d's receive slot of source r, written over NVLink by the tile kernel itself
    int32_t* dstP = d == r ? s->precv + (size_t)r * s->blk
                           : (s->p2p ? s->peer_recv[d] + (size_t)r * s->blk : s->psend + (size_t)d * s->blk);
    rc = edt_stage_zy(st, s->rec, s->nxl, s->ny, s->NW, d * s->wl, s->wl, dstP, 32, (int64_t)s->ny * s->nxl * 32,
                      (int64_t)s->nxl * 32);
    if (rc) return fuel_fail(nullptr, rc, "zy stage failed");
    if (d == r || s->p2p) continue;
    FUEL_CUDA(nullptr, cudaEventRecord(s->ev_round[k], st));
    FUEL_CUDA(nullptr, cudaStreamWaitEvent(s->comm_stream, s->ev_round[k], 0));
    FUEL_NCCL(g_nccl.GroupStart());
    FUEL_NCCL(g_nccl.Send(s->psend + (size_t)d * s->blk, s->blk, ncclInt32, d, c->comm, s->comm_stream));
    FUEL_NCCL(g_nccl.Recv(s->precv + (size_t)src * s->blk, s->blk, ncclInt32, src, c->comm, s->comm_stream));
    FUEL_NCCL(g_nccl.GroupEnd());
  }
  FUEL_CUDA(nullptr, cudaEventRecord(s->ev_t[2], st));
  if (s->p2p) {
    // every source has finished storing into my receive buffer once this tiny all-gather completes (a kernel's
    // stores are performed at system scope when it ends; the collective orders the ranks' streams).  The slots of
    // the NEXT update are not written before step 1 of that update has met every peer, i.e. after its x tiles.
    FUEL_NCCL(g_nccl.AllGather(s->sync_buf, s->sync_buf + 1, 1, ncclInt32, c->comm, st));
  } else {
    FUEL_CUDA(nullptr, cudaEventRecord(s->ev_comm_done, s->comm_stream));
    FUEL_CUDA(nullptr, cudaStreamWaitEvent(st, s->ev_comm_done, 0));
  }
  FUEL_CUDA(nullptr, cudaEventRecord(s->ev_t[3], st));
  // 4. x tiles on my z-slab
  rc = edt_stage_x(st, s->precv, (int64_t)s->nxl * 32, (int64_t)s->ny * s->nxl * 32, (int64_t)s->blk, s->nxl, s->nx, s->ny,
                   s->wl, (float*)dist_slab_dev, s->nzl, 32, (int64_t)s->ny * s->nzl, s->nzl, (float)s->res, 0);
  if (rc) return fuel_fail(nullptr, rc, "x stage failed");
  FUEL_CUDA(nullptr, cudaEventRecord(s->ev_t[4], st));
  s->timed = true;
  return 0;
}

// device times of the last update on this rank: [0] occupancy exchange, [1] zpack + zy tiles (the partial's
// exchange rounds run beside them), [2] wait for the last exchange rounds, [3] x tiles, [4] whole update
int fuelgpu_sharded_esdf_last_timing(FuelShardedEsdf* s, float ms[5]) {
  if (!s || !ms) return fuel_fail(nullptr, FUELGPU_EINVAL, "null argument");
  for (int i = 0; i < 5; ++i) ms[i] = -1.f;
  if (!s->timed) return 0;
  FUEL_CUDA(nullptr, cudaEventSynchronize(s->ev_t[4]));
  for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&ms[i], s->ev_t[i], s->ev_t[i + 1]);
  cudaEventElapsedTime(&ms[4], s->ev_t[0], s->ev_t[4]);
  return 0;
}

int64_t fuelgpu_sharded_esdf_bytes_exchanged(const FuelShardedEsdf* s) {
  if (!s) return 0;
  const int64_t G = s->c->nranks;
  const int64_t occ = (int64_t)s->nxl * s->ny * s->nzl * (G - 1);
  return occ + (int64_t)s->blk * 4 * (G - 1);  // sent by this rank per update
}

// all-gather of the z-slabs of distance_buffer_ (what a trajectory batch on every rank samples): out is
// [G][nx][ny][nzl] float32, slab g = rank g's planes
int fuelgpu_sharded_esdf_allgather(FuelShardedEsdf* s, void* cuda_stream, const void* dist_slab_dev, void* out_dev) {
  if (!s || !dist_slab_dev || !out_dev) return fuel_fail(nullptr, FUELGPU_EINVAL, "null argument");
  const size_t cnt = (size_t)s->nx * s->ny * s->nzl;
  FUEL_NCCL(g_nccl.AllGather(dist_slab_dev, out_dev, cnt, ncclFloat, s->c->comm, (cudaStream_t)cuda_stream));
  return 0;
}

}  // extern "C"
