// Library-side multi-GPU: a communicator over the B200s of one box and the sharded Galerkin assembly with its ONE
// collective (north_star: "Q assembly shards naturally by output mode i ... with a single NCCL
// all-gather over NVLink so that every GPU holds the full Q").
//
// Two ways to obtain a communicator, so that neither a Julia process nor a one-process-per-GPU launcher needs an NCCL
// binding of its own:
//   ca_comm_init_all(ngpus)            one process, one host thread drives all devices (ncclCommInitAll)
//   ca_comm_unique_id + _init_rank     one process per GPU; the 128-byte id travels by whatever the launcher has
// NCCL is resolved at run time (dlopen "libnccl.so.2": the copy torch already mapped, or the system one), so the
// library has no link-time dependency on it and single-GPU users never load it.
//
// ca_assemble_sharded: every rank assembles a contiguous range of output modes (ca::assemble_slab), the slab sizes are
// exchanged (one 8-byte all-gather), and every piece of every slab -- four matrix row blocks, three index columns, the
// values -- is broadcast straight into its place in the full arrays, all inside ONE ncclGroup (an all-gather-v): a single
// fused collective per device, no padding to the largest slab, no staging, no concatenation pass.  Rank order reproduces the reference's
// lexicographic (i,j,k) order (SparseBilinear.jl:24-49) because ranks own increasing ranges of i.
#include <dlfcn.h>

#include <algorithm>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "assembly.h"
#include "comm.h"
#include "model.h"

using namespace ca;

namespace ca {

const NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      err = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "");
      return;
    }
    auto sym = [&](const char* name) {
      void* p = dlsym(api.handle, name);
      if (!p) err = std::string("NCCL symbol missing: ") + name;
      return p;
    };
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  });
  if (!err.empty()) fail(CA_ERR_UNSUPPORTED, "%s", err.c_str());
  return api;
}

}  // namespace ca

ca_comm::~ca_comm() {
  for (size_t d = 0; d < comms.size(); ++d) {
    cudaSetDevice(devices[d]);
    if (comms[d] && nccl().CommDestroy) nccl().CommDestroy(comms[d]);
    if (streams[d]) cudaStreamDestroy(streams[d]);
  }
}

namespace {

std::unique_ptr<ca_comm> g_default_comm;  // ca_init / ca_shutdown

// contiguous ranges of output modes per rank; `align`: boundaries on multiples of it (the dense path computes whole
// 128-row output tiles); trailing ranks may be empty.  Same rule as the Python mirror's row_partition.
std::vector<int64_t> row_partition(int64_t m, int world, int64_t align) {
  const int64_t units = (m + align - 1) / align, base = units / world, extra = units % world;
  std::vector<int64_t> bounds{0};
  for (int r = 0; r < world; ++r) bounds.push_back(std::min(m, bounds.back() + (base + (r < extra ? 1 : 0)) * align));
  bounds.back() = m;
  return bounds;
}

// runs fn(d) for every local device, on its own host thread when there are several (the slab assemblies and model
// builds of different devices are independent and each synchronises its own device)
void for_each_device(const ca_comm& c, const std::function<void(size_t)>& fn) {
  const size_t n = c.devices.size();
  if (n == 1) {
    CA_CUDA(cudaSetDevice(c.devices[0]));
    fn(0);
    return;
  }
  std::vector<std::string> errors(n);
  std::vector<int32_t> codes(n, CA_OK);
  std::vector<std::thread> pool;
  for (size_t d = 0; d < n; ++d)
    pool.emplace_back([&, d] {
      try {
        CA_CUDA(cudaSetDevice(c.devices[d]));
        fn(d);
      } catch (const Error& e) {
        codes[d] = e.code;
        errors[d] = e.what();
      } catch (const std::exception& e) {
        codes[d] = CA_ERR_INVALID;
        errors[d] = e.what();
      }
    });
  for (auto& t : pool) t.join();
  for (size_t d = 0; d < n; ++d)
    if (codes[d] != CA_OK) fail(codes[d], "device %d: %s", c.devices[d], errors[d].c_str());
}

}  // namespace

ca_comm* default_comm() { return g_default_comm.get(); }

extern "C" {

int32_t ca_comm_init_all(int32_t ngpus, const int32_t* devices, ca_comm** out) {
  return guarded([&] {
    if (!out) fail(CA_ERR_INVALID, "ca_comm_init_all: out is null");
    *out = nullptr;
    require_device();
    int ndev = 0;
    CA_CUDA(cudaGetDeviceCount(&ndev));
    if (ngpus < 1 || ngpus > ndev) fail(CA_ERR_INVALID, "ca_comm_init_all: %d GPUs requested, %d visible", ngpus, ndev);
    int prev = 0;
    CA_CUDA(cudaGetDevice(&prev));
    std::unique_ptr<ca_comm> c(new ca_comm);
    c->nranks = ngpus;
    for (int d = 0; d < ngpus; ++d) {
      c->devices.push_back(devices ? devices[d] : d);
      c->ranks.push_back(d);
    }
    c->comms.assign((size_t)ngpus, nullptr);
    c->streams.assign((size_t)ngpus, nullptr);
    if (ngpus > 1) CA_NCCL(nccl().CommInitAll(c->comms.data(), ngpus, c->devices.data()));
    for (int d = 0; d < ngpus; ++d) {
      CA_CUDA(cudaSetDevice(c->devices[d]));
      CA_CUDA(cudaStreamCreateWithFlags(&c->streams[d], cudaStreamNonBlocking));
    }
    CA_CUDA(cudaSetDevice(prev));
    *out = c.release();
  });
}

int32_t ca_comm_unique_id(char* id128) {
  return guarded([&] {
    if (!id128) fail(CA_ERR_INVALID, "ca_comm_unique_id: null argument");
    ncclUniqueId id;
    CA_NCCL(nccl().GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "NCCL unique id is 128 bytes");
    memcpy(id128, &id, 128);
  });
}

int32_t ca_comm_init_rank(int32_t nranks, int32_t rank, const char* id128, ca_comm** out) {
  return guarded([&] {
    if (!out || !id128) fail(CA_ERR_INVALID, "ca_comm_init_rank: null argument");
    *out = nullptr;
    if (nranks < 1 || rank < 0 || rank >= nranks) fail(CA_ERR_INVALID, "ca_comm_init_rank: rank %d of %d", rank, nranks);
    require_device();
    std::unique_ptr<ca_comm> c(new ca_comm);
    int dev = 0;
    CA_CUDA(cudaGetDevice(&dev));
    c->nranks = nranks;
    c->devices.push_back(dev);
    c->ranks.push_back(rank);
    c->comms.assign(1, nullptr);
    c->streams.assign(1, nullptr);
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    if (nranks > 1) CA_NCCL(nccl().CommInitRank(&c->comms[0], nranks, id, rank));
    CA_CUDA(cudaStreamCreateWithFlags(&c->streams[0], cudaStreamNonBlocking));
    *out = c.release();
  });
}

int32_t ca_comm_destroy(ca_comm* c) {
  return guarded([&] { delete c; });
}

int32_t ca_comm_info(const ca_comm* c, int32_t* nranks, int32_t* nlocal, int32_t* first_rank, int32_t* nccl_version) {
  return guarded([&] {
    if (!c) fail(CA_ERR_INVALID, "null communicator");
    if (nranks) *nranks = c->nranks;
    if (nlocal) *nlocal = (int32_t)c->devices.size();
    if (first_rank) *first_rank = c->ranks[0];
    if (nccl_version) {
      int v = 0;
      if (c->nranks > 1) CA_NCCL(nccl().GetVersion(&v));
      *nccl_version = v;
    }
  });
}

int32_t ca_init(int32_t ngpus) {
  return guarded([&] {
    ca_comm* c = nullptr;
    const int32_t st = ca_comm_init_all(ngpus, nullptr, &c);
    if (st != CA_OK) fail(st, "%s", ca_last_error());
    g_default_comm.reset(c);
  });
}

int32_t ca_shutdown(void) {
  return guarded([&] { g_default_comm.reset(); });
}

int32_t ca_assemble_sharded(ca_comm* comm, double alpha, double gamma, const int64_t* ijkl, int64_t m,
                            int32_t normalize, const double* mode_scale, int32_t dense, int32_t nx, int32_t ny,
                            int32_t nz, ca_assembly** out) {
  return guarded([&] {
    ca_comm* c = comm ? comm : g_default_comm.get();
    if (!c) fail(CA_ERR_INVALID, "ca_assemble_sharded: no communicator (ca_init / ca_comm_init_*)");
    if (!out) fail(CA_ERR_INVALID, "ca_assemble_sharded: out is null");
    const size_t nloc = c->devices.size();
    for (size_t d = 0; d < nloc; ++d) out[d] = nullptr;
    const int world = c->nranks;
    const std::vector<int64_t> bounds = row_partition(m, world, dense ? 128 : 1);
    int prev = 0;
    CA_CUDA(cudaGetDevice(&prev));
    AssembleOpts opts;
    opts.dense = dense != 0;
    opts.nx = nx;
    opts.ny = ny;
    opts.nz = nz;
    opts.mode_scale = mode_scale;

    // ---- every rank its slab
    std::vector<std::unique_ptr<ca_assembly>> slab(nloc), full(nloc);
    for_each_device(*c, [&](size_t d) {
      ca_assembly* s = nullptr;
      const int r = c->ranks[d];
      const int32_t st = assemble_slab(alpha, gamma, ijkl, m, normalize, bounds[(size_t)r], bounds[(size_t)r + 1], opts, &s);
      if (st != CA_OK) fail(st, "%s", ca_last_error());
      slab[d].reset(s);
    });
    if (world == 1) {  // nothing to exchange
      out[0] = slab[0].release();
      CA_CUDA(cudaSetDevice(prev));
      return;
    }

    // ---- slab sizes: one 8-byte all-gather
    std::vector<DeviceBuffer<int64_t>> dsz(nloc), dall(nloc);
    for (size_t d = 0; d < nloc; ++d) {
      CA_CUDA(cudaSetDevice(c->devices[d]));
      dsz[d].upload(&slab[d]->nnz, 1, c->streams[d]);
      dall[d].alloc((size_t)world);
    }
    CA_NCCL(nccl().GroupStart());
    for (size_t d = 0; d < nloc; ++d)
      CA_NCCL(nccl().AllGather(dsz[d].ptr, dall[d].ptr, 1, ncclInt64, c->comms[d], c->streams[d]));
    CA_NCCL(nccl().GroupEnd());
    std::vector<int64_t> nnz_of((size_t)world), noff((size_t)world + 1, 0);
    CA_CUDA(cudaSetDevice(c->devices[0]));
    CA_CUDA(cudaMemcpyAsync(nnz_of.data(), dall[0].ptr, (size_t)world * sizeof(int64_t), cudaMemcpyDeviceToHost, c->streams[0]));
    for (size_t d = 0; d < nloc; ++d) {
      CA_CUDA(cudaSetDevice(c->devices[d]));
      CA_CUDA(cudaStreamSynchronize(c->streams[d]));
    }
    for (int r = 0; r < world; ++r) noff[(size_t)r + 1] = noff[(size_t)r] + nnz_of[(size_t)r];
    const int64_t nnz = noff[(size_t)world];

    // ---- how the COO arrays travel (the matrices always go as broadcasts):
    //   broadcasts  every piece of every slab broadcast in place (no padding, no extra pass)              -- the default
    //   padded      slabs padded to the largest one, ONE in-place ncclAllGather per array, then a device-to-device
    //               compaction into the contiguous arrays                                      -- CA_GATHER=allgather
    // Measured at m = 2962 (9.48 GB), both in one run: 2 GPUs 8.2 ms (broadcasts) / 12.4 ms (padded); 8 GPUs 28.8 ms
    // (288 GB/s bus bandwidth) / 31.1 ms -- NCCL's own all-gather is no faster than the 64 grouped broadcasts on this
    // node, and the compaction comes on top.
    bool padded = false;
    if (const char* e = getenv("CA_GATHER")) padded = e[0] == 'a';
    size_t nmax = 0;
    for (int r = 0; r < world; ++r) nmax = std::max(nmax, (size_t)nnz_of[(size_t)r]);
    if (nmax == 0) padded = false;
    std::vector<DeviceBuffer<int64_t>> pad_i(nloc), pad_j(nloc), pad_k(nloc);
    std::vector<DeviceBuffer<double>> pad_v(nloc);

    // ---- full arrays on every device, own slab copied into place
    std::vector<cudaEvent_t> ev0(nloc), ev1(nloc);
    for (size_t d = 0; d < nloc; ++d) {
      CA_CUDA(cudaSetDevice(c->devices[d]));
      full[d].reset(new ca_assembly);
      ca_assembly& F = *full[d];
      const ca_assembly& S = *slab[d];
      F.device = c->devices[d];
      F.m = m;
      F.r0 = 0;
      F.r1 = m;
      F.nnz = nnz;
      F.seconds = S.seconds;
      F.contract_seconds = S.contract_seconds;
      F.contract_flops = S.contract_flops;
      for (int q = 0; q < 3; ++q) F.grid[q] = S.grid[q];
      F.shear = S.shear;
      F.world = world;
      DeviceBuffer<double>* fm[4] = {&F.B, &F.A1, &F.A2, &F.D};
      const DeviceBuffer<double>* sm[4] = {&S.B, &S.A1, &S.A2, &S.D};
      const int r = c->ranks[d];
      const size_t rows = (size_t)(bounds[(size_t)r + 1] - bounds[(size_t)r]);
      for (int q = 0; q < 4; ++q) {
        fm[q]->alloc((size_t)m * m);
        if (rows)
          CA_CUDA(cudaMemcpyAsync(fm[q]->ptr + (size_t)bounds[(size_t)r] * m, sm[q]->ptr, rows * (size_t)m * sizeof(double),
                                  cudaMemcpyDeviceToDevice, c->streams[d]));
      }
      F.oi.alloc((size_t)nnz);
      F.oj.alloc((size_t)nnz);
      F.ok.alloc((size_t)nnz);
      F.oval.alloc((size_t)nnz);
      const size_t n = (size_t)S.nnz;
      if (padded) {  // own slab into its slot of the staging arrays
        pad_i[d].alloc((size_t)world * nmax);
        pad_j[d].alloc((size_t)world * nmax);
        pad_k[d].alloc((size_t)world * nmax);
        pad_v[d].alloc((size_t)world * nmax);
      }
      const size_t at = padded ? (size_t)r * nmax : (size_t)noff[(size_t)r];
      if (n) {
        CA_CUDA(cudaMemcpyAsync((padded ? pad_i[d].ptr : F.oi.ptr) + at, S.oi.ptr, n * 8, cudaMemcpyDeviceToDevice, c->streams[d]));
        CA_CUDA(cudaMemcpyAsync((padded ? pad_j[d].ptr : F.oj.ptr) + at, S.oj.ptr, n * 8, cudaMemcpyDeviceToDevice, c->streams[d]));
        CA_CUDA(cudaMemcpyAsync((padded ? pad_k[d].ptr : F.ok.ptr) + at, S.ok.ptr, n * 8, cudaMemcpyDeviceToDevice, c->streams[d]));
        CA_CUDA(cudaMemcpyAsync((padded ? pad_v[d].ptr : F.oval.ptr) + at, S.oval.ptr, n * 8, cudaMemcpyDeviceToDevice, c->streams[d]));
      }
      CA_CUDA(cudaEventCreate(&ev0[d]));
      CA_CUDA(cudaEventCreate(&ev1[d]));
      CA_CUDA(cudaEventRecord(ev0[d], c->streams[d]));
    }

    // ---- the all-gather(-v), one group.  (Also built and measured: the same pieces as grouped point-to-point
    // ncclSend / ncclRecv, all pairs at once -- equal at 2 GPUs, 40.2 ms at 8 and 6.9 s of connection set-up for the 56
    // pairs on first use.)
    CA_NCCL(nccl().GroupStart());
    for (size_t d = 0; d < nloc; ++d) {
      ca_assembly& F = *full[d];
      double* fm[4] = {F.B.ptr, F.A1.ptr, F.A2.ptr, F.D.ptr};
      for (int r = 0; r < world; ++r) {
        const size_t rows = (size_t)(bounds[(size_t)r + 1] - bounds[(size_t)r]), n = (size_t)nnz_of[(size_t)r];
        const size_t at = (size_t)noff[(size_t)r];
        if (rows)
          for (int q = 0; q < 4; ++q) {
            double* p = fm[q] + (size_t)bounds[(size_t)r] * m;
            CA_NCCL(nccl().Broadcast(p, p, rows * (size_t)m, ncclFloat64, r, c->comms[d], c->streams[d]));
          }
        if (n && !padded) {
          CA_NCCL(nccl().Broadcast(F.oi.ptr + at, F.oi.ptr + at, n, ncclInt64, r, c->comms[d], c->streams[d]));
          CA_NCCL(nccl().Broadcast(F.oj.ptr + at, F.oj.ptr + at, n, ncclInt64, r, c->comms[d], c->streams[d]));
          CA_NCCL(nccl().Broadcast(F.ok.ptr + at, F.ok.ptr + at, n, ncclInt64, r, c->comms[d], c->streams[d]));
          CA_NCCL(nccl().Broadcast(F.oval.ptr + at, F.oval.ptr + at, n, ncclFloat64, r, c->comms[d], c->streams[d]));
        }
      }
      if (padded) {  // in place: the send buffer is this rank's slot of the receive buffer
        const size_t mine = (size_t)c->ranks[d] * nmax;
        CA_NCCL(nccl().AllGather(pad_i[d].ptr + mine, pad_i[d].ptr, nmax, ncclInt64, c->comms[d], c->streams[d]));
        CA_NCCL(nccl().AllGather(pad_j[d].ptr + mine, pad_j[d].ptr, nmax, ncclInt64, c->comms[d], c->streams[d]));
        CA_NCCL(nccl().AllGather(pad_k[d].ptr + mine, pad_k[d].ptr, nmax, ncclInt64, c->comms[d], c->streams[d]));
        CA_NCCL(nccl().AllGather(pad_v[d].ptr + mine, pad_v[d].ptr, nmax, ncclFloat64, c->comms[d], c->streams[d]));
      }
    }
    CA_NCCL(nccl().GroupEnd());
    for (size_t d = 0; d < nloc; ++d) {
      CA_CUDA(cudaSetDevice(c->devices[d]));
      if (padded) {  // compaction: slot r -> its place in the contiguous arrays
        ca_assembly& F = *full[d];
        for (int r = 0; r < world; ++r) {
          const size_t n = (size_t)nnz_of[(size_t)r], from = (size_t)r * nmax, to = (size_t)noff[(size_t)r];
          if (!n) continue;
          CA_CUDA(cudaMemcpyAsync(F.oi.ptr + to, pad_i[d].ptr + from, n * 8, cudaMemcpyDeviceToDevice, c->streams[d]));
          CA_CUDA(cudaMemcpyAsync(F.oj.ptr + to, pad_j[d].ptr + from, n * 8, cudaMemcpyDeviceToDevice, c->streams[d]));
          CA_CUDA(cudaMemcpyAsync(F.ok.ptr + to, pad_k[d].ptr + from, n * 8, cudaMemcpyDeviceToDevice, c->streams[d]));
          CA_CUDA(cudaMemcpyAsync(F.oval.ptr + to, pad_v[d].ptr + from, n * 8, cudaMemcpyDeviceToDevice, c->streams[d]));
        }
      }
      CA_CUDA(cudaEventRecord(ev1[d], c->streams[d]));
    }
    for (size_t d = 0; d < nloc; ++d) {
      CA_CUDA(cudaSetDevice(c->devices[d]));
      CA_CUDA(cudaStreamSynchronize(c->streams[d]));
      float ms = 0;
      CA_CUDA(cudaEventElapsedTime(&ms, ev0[d], ev1[d]));
      full[d]->gather_seconds = ms * 1e-3;
      full[d]->gather_bytes = (int64_t)4 * m * m * 8 + nnz * 32;
      cudaEventDestroy(ev0[d]);
      cudaEventDestroy(ev1[d]);
    }
    for (size_t d = 0; d < nloc; ++d) out[d] = full[d].release();
    CA_CUDA(cudaSetDevice(prev));
  });
}

int32_t ca_assembly_gather_info(const ca_assembly* h, int32_t* world, double* gather_seconds, int64_t* gather_bytes) {
  return guarded([&] {
    if (!h) fail(CA_ERR_INVALID, "null handle");
    if (world) *world = h->world;
    if (gather_seconds) *gather_seconds = h->gather_seconds;
    if (gather_bytes) *gather_bytes = h->gather_bytes;
  });
}

}  // extern "C"
