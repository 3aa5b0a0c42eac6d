/*
 * engine.cu -- device side of the path: unified device layout of all hosted domains, the
 * Green-Gauss tile kernel (sm_100a), device pack / unpack, and the halo exchange pipeline.
 *
 * Reference functions replaced:
 *   private_compute_gradients_gg ............ src/gradients.c:25-147   -> gg_tile_pipe_kernel (gg_kernels.cuh); gg_tile_kernel = second implementation
 *   compute_gradients_gg_<variant> .......... src/gradients.c:150-335  -> run_iteration()
 *   private_get_color_and_exchange .......... src/rangelist.c:838-889  -> stream/event pipeline
 *   initiate_thread_comm_mpi_send / _pack ... src/threads.c:187-346    -> boundary tiles first; pack fused into the kernel (export lists);
 *                                                                        gaspi_async on several GPUs: direct stores into peer memory
 *   exchange_dbl_copy_in / copy_out ......... src/threads.c:791-869    -> fused pack / rows_copy_kernel (unpack, permutations)
 *   exchange_dbl_mpi_send/_post_recv/_bulk_sync/_early_recv/_async
 *                                             src/exchange_data_mpi.c:96-543 -> grouped ncclSend/ncclRecv per peer GPU
 *   exchange_dbl_gaspi_write / mpidma_write . src/exchange_data_gaspi.c:105-151, exchange_data_mpidma.c:93-127
 *                                                                     -> CUDA-IPC put + notify, or direct halo stores (run_iteration_direct)
 *   init_threads ............................ src/threads.c:730-788    -> registers the domain; cfdp_commit builds the schedule
 *
 * Layout in HBM (one process = one GPU, all hosted domains concatenated):
 *   var  [rows][7]  f64   0.5 * var; rows = for each domain: tiles (boundary tiles first, each padded to 16 rows), then ghosts
 *   hhalo           per tile: the var rows of its halo points, contiguous (halo_pack_kernel, rebuilt when var changes)
 *   grad [rows][21] f64
 *   pvol [rows]     f64
 *   blob            tile blobs (normals, halo row list, ELL adjacency), see common.h
 *   tiles           TileDesc list: boundary tiles of all domains, then interior tiles of all domains
 */
#include <cuda_runtime.h>
#include <cuda.h>
#include <dlfcn.h>
#include <string.h>
#include <unistd.h>
#include <errno.h>
#include <fcntl.h>
#include <time.h>
#include <sys/stat.h>
#include <algorithm>
#include <map>
#include <mutex>
#include <vector>
#include <omp.h>
#include "common.h"
#include "gg_kernels.cuh"
#include "flux_kernels.cuh"

#define CUDA_CHECK(call)                                                                           \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      fprintf(stderr, "Error: '%s' [%s:%i]: %s\n", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(EXIT_FAILURE);                                                                          \
    }                                                                                              \
  } while (0)

/* ------------------------------------------------------------------------------------------
 * NCCL, resolved at run time (the process usually has torch's bundled libnccl.so.2 loaded)
 * ---------------------------------------------------------------------------------------- */
typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm;
enum { NCCL_INT32 = 2, NCCL_FLOAT64 = 8 };
struct NcclApi {
  void *h = nullptr;
  int (*GetUniqueId)(nccl_uid *) = nullptr;
  int (*CommInitRank)(nccl_comm *, int, nccl_uid, int) = nullptr;
  int (*CommDestroy)(nccl_comm) = nullptr;
  int (*Send)(const void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
  int (*GroupStart)(void) = nullptr;
  int (*GroupEnd)(void) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static void nccl_load(void)
{
  if (g_nccl.h) return;
  const char *cands[] = { getenv("CFDP_NCCL_LIB"), "libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2" };
  for (const char *c : cands) {
    if (!c || !*c) continue;
    g_nccl.h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) { fprintf(stderr, "Error: cannot load libnccl.so.2 (set CFDP_NCCL_LIB): %s\n", dlerror()); exit(EXIT_FAILURE); }
#define SYM(field, name) do { *(void **)(&g_nccl.field) = dlsym(g_nccl.h, name); ASSERT(g_nccl.field != nullptr); } while (0)
  SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
  SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv"); SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
}
#define NCCL_CHECK(call)                                                                            \
  do {                                                                                              \
    int r_ = (call);                                                                                \
    if (r_ != 0) {                                                                                  \
      fprintf(stderr, "Error: '%s' [%s:%i]: %s\n", #call, __FILE__, __LINE__, g_nccl.GetErrorString(r_)); \
      exit(EXIT_FAILURE);                                                                           \
    }                                                                                               \
  } while (0)

/* ------------------------------------------------------------------------------------------
 * Engine state
 * ---------------------------------------------------------------------------------------- */
struct PeerPlan {
  int proc = -1;
  long long send_off = 0, send_rows = 0; /* row offsets into the packed send / recv buffers */
  long long recv_off = 0, recv_rows = 0;
  /* one-sided backend (CUDA IPC): where my rows go in the peer's receive buffer, which of the peer's arrival
   * counters is mine, and the peer's buffers mapped into this process */
  long long remote_recv_off = 0, remote_recv_total = 0;
  int remote_slot = 0;
  double *peer_recvbuf = nullptr;
  unsigned long long *peer_flags = nullptr;
  /* direct halo stores: the peer's grad array mapped into this process, and for every row of my packed send
   * segment the row of the peer's grad it belongs to (the peer's unpack list, exchanged at setup) */
  double *peer_grad = nullptr;
  std::vector<uint32_t> peer_rows;
  bool opened = false;   /* the three mappings above came from cudaIpcOpenMemHandle (not loopback aliases) */
};
/* slots of the arrival-counter array every rank owns (d_arrived, 256 x u64), by peer index i:
 *   [i]        stage number of the last put into my receive window            (put + notify variants)
 *   [64 + i]   rows peer i has stored into my ghost rows, cumulative           (direct halo stores)
 *   [128 + i]  epoch up to which peer i has consumed the ghost rows I stored   (direct halo stores: write credit) */
enum { FLAG_STAGE = 0, FLAG_ROWS = 64, FLAG_CREDIT = 128, MAX_PEERS = 8 };

struct Engine {
  bool configured = false, planned = false, committed = false, have_device = false;
  int proc_rank = 0, nprocs = 0, ndomains_total = 0, per_proc = 0, device = 0;
  int resident = 0, exact = 1;
  std::vector<Domain *> doms;  /* hosted, by ascending domain id */
  std::mutex mu;
  /* options */
  ScheduleOptions sopt;
  /* device */
  long long rows = 0, ntiles = 0, nbtiles = 0;
  double *d_var = nullptr, *d_grad = nullptr, *d_pvol = nullptr;
  double *d_hhalo = nullptr; long long halo_rows = 0; /* packed halo rows: per tile a contiguous copy of the half-var rows of its halo positions (gg_kernels.cuh) */
  unsigned char *d_fblob = nullptr; TileDesc *d_ftiles = nullptr; std::vector<TileDesc> h_ftiles; std::vector<size_t> fblob_base; size_t fblob_bytes = 0; /* pseudo-flux blobs (same tile order as h_tiles) */
  double *d_flux = nullptr; int with_flux = 0; uint32_t flux_smem = 0; double last_flux_ms = 0; long long flux_alg_bytes = 0; /* pseudo flux (flux.c), lazily allocated */
  unsigned char *d_blob = nullptr;
  TileDesc *d_tiles = nullptr;
  size_t blob_bytes = 0;
  int smem_bytes = 0, region0_doubles = 0, block_threads = 0, ctas_per_sm = 2, var_refresh = 0;
  int kernel_version = 2, chunk = 16, smem_v1 = 0, persistent = 0; ggk::PipeLayout pipe = {}; uint32_t max_hvpv = 0;
  unsigned long long *d_progress = nullptr; unsigned long long progress_target = 0; int fused_signal = 1;
  /* fused pack: per boundary tile, the rows other domains need (export lists), written by the gradient kernel itself */
  std::vector<uint32_t> h_exp_off, h_exp_src, h_exp_dst; uint32_t *d_exp_off = nullptr, *d_exp_src = nullptr, *d_exp_dst = nullptr; int fused_pack = 1;
  /* one-sided exchange: double-buffered receive window + per-peer arrival counters (exchange_data_gaspi.c:105-151) */
  bool loopback = false;           /* test mode (CFDP_LOOPBACK=1): halo rows between domains of this GPU take the inter-GPU path, this process being its own peer */
  bool direct_ready = false; unsigned long long direct_epoch = 0; uint32_t *d_exp_src_direct = nullptr, *d_exp_dst_direct = nullptr, *d_sig_off = nullptr, *d_sig_ent = nullptr;
  int last_transport = 0;          /* what the last exchange used: 0 none, 1 on-device copies only, 2 NCCL send/recv, 3 IPC put + notify, 4 direct stores into peer memory */
  bool ipc_ready = false; double *d_recvwin = nullptr; unsigned long long *d_arrived = nullptr; unsigned long long ipc_stage = 0; uint32_t max_footprint = 0;
  size_t max_blob = 0; int max_nhalo = 0;
  std::vector<int *> d_rowmap;     /* per hosted domain: [nall] global device row of host point */
  double *d_stage = nullptr; size_t stage_bytes = 0;
  /* exchange plan */
  long long n_local = 0;            /* ghost rows filled from a domain hosted on this GPU */
  uint32_t *d_loc_dst = nullptr, *d_loc_src = nullptr;
  long long n_send = 0, n_recv = 0; /* rows packed for / unpacked from other GPUs */
  uint32_t *d_send_rows = nullptr, *d_recv_rows = nullptr;
  double *d_sendbuf = nullptr, *d_recvbuf = nullptr;
  std::vector<PeerPlan> peers;
  /* host copies of the plan (built by cfdp_plan without touching the device) */
  std::vector<TileDesc> h_tiles;
  std::vector<size_t> blob_base;
  std::vector<uint32_t> h_loc_dst, h_loc_src, h_send_rows, h_recv_rows;
  int max_nfaces = 0, max_nloc = 0, max_npts = 0; size_t max_stage = 0;
  std::map<std::pair<int, int>, std::vector<uint32_t>> send_rows_of, recv_rows_of; /* (domain, partner) -> device rows */
  std::vector<std::vector<int>> point_of_row; /* per hosted domain: domain-relative row -> host point (-1 padding) */
  cudaStream_t s_comp = nullptr, s_comm = nullptr;
  cudaEvent_t ev_b = nullptr, ev_x = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  cudaEvent_t timeline_ek = nullptr, timeline_ex = nullptr;
  nccl_comm comm = nullptr;
  /* stats */
  long long launches = 0;
  double last_kernel_ms = 0;
  long long nfaces = 0, nown = 0, nall = 0, tile_faces = 0, halo_refs = 0, alg_bytes = 0;
};
static Engine g_eng;

Engine *engine_get(void) { return &g_eng; }
int engine_proc_of_domain(int domain) { return g_eng.per_proc > 0 ? domain / g_eng.per_proc : 0; }
int engine_num_hosted(void) { return (int)g_eng.doms.size(); }
Domain *engine_hosted(int i) { return g_eng.doms[(size_t)i]; }
Domain *engine_domain_by_id(int id) { for (Domain *d : g_eng.doms) if (d->id == id) return d; return nullptr; }
Domain *engine_find_domain(const void *p) { for (Domain *d : g_eng.doms) if ((const void *)d->cd == p || (const void *)d->sd == p) return d; return nullptr; }

Domain *engine_register_domain(comm_data *cd, int id)
{
  std::lock_guard<std::mutex> lk(g_eng.mu);
  ASSERT(!g_eng.planned);
  Domain *d = engine_domain_by_id(id);
  if (!d) {
    d = new Domain(); d->id = id; g_eng.doms.push_back(d);
    std::sort(g_eng.doms.begin(), g_eng.doms.end(), [](Domain *a, Domain *b) { return a->id < b->id; });
  }
  d->cd = cd; d->sd = nullptr; d->comm_read = d->tables_done = d->threads_inited = false;
  return d;
}

static void ensure_device(void)
{
  Engine &E = g_eng;
  if (E.have_device) return;
  int n = 0;
  cudaError_t err = cudaGetDeviceCount(&n);
  if (err != cudaSuccess || n == 0) {
    fprintf(stderr, "Error: no CUDA device: the Green-Gauss path has no CPU fallback [%s:%i]\n", __FILE__, __LINE__);
    exit(EXIT_FAILURE);
  }
  ASSERT(E.device < n);
  CUDA_CHECK(cudaSetDevice(E.device));
  int prio_lo = 0, prio_hi = 0;
  CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  CUDA_CHECK(cudaStreamCreateWithPriority(&E.s_comp, cudaStreamNonBlocking, prio_lo));
  /* the exchange kernels (pack, NCCL, unpack) must get the SM slots the gradient grid frees: highest priority */
  CUDA_CHECK(cudaStreamCreateWithPriority(&E.s_comm, cudaStreamNonBlocking, prio_hi));
  CUDA_CHECK(cudaEventCreateWithFlags(&E.ev_b, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreateWithFlags(&E.ev_x, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreate(&E.ev_t0));
  CUDA_CHECK(cudaEventCreate(&E.ev_t1));
  E.have_device = true;
}

static bool device_present(void)
{
  static int state = -1;
  if (state < 0) { int n = 0; state = (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) ? 1 : 0; if (!state) (void)cudaGetLastError(); }
  return state == 1;
}

void *engine_alloc_pinned(size_t bytes)
{
  ASSERT(bytes > 0);
  void *p = nullptr;
  if (device_present() && !getenv("CFDP_NO_PINNED")) {
    if (g_eng.configured) (void)cudaSetDevice(g_eng.device);
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess) return p;
    (void)cudaGetLastError();
  }
  p = malloc(bytes); /* host container only (e.g. loader tests on a machine without a GPU) */
  ASSERT(p != NULL);
  return p;
}
void engine_free_pinned(void *p) { if (p && cudaFreeHost(p) != cudaSuccess) { (void)cudaGetLastError(); free(p); } }

static int env_int(const char *name, int dflt) { const char *s = getenv(name); return (s && *s) ? atoi(s) : dflt; }

extern "C" int cfdp_configure(int proc_rank, int nprocs, int ndomains_total, int device)
{
  Engine &E = g_eng;
  if (E.planned || !E.doms.empty()) return -1;
  if (nprocs < 1 || proc_rank < 0 || proc_rank >= nprocs || ndomains_total < nprocs || ndomains_total % nprocs) return -2;
  E.proc_rank = proc_rank; E.nprocs = nprocs; E.ndomains_total = ndomains_total; E.per_proc = ndomains_total / nprocs;
  E.device = device >= 0 ? device : env_int("LOCAL_RANK", 0);
  E.sopt.tile_points = env_int("CFDP_TILE_POINTS", 256);
  E.sopt.max_faces = env_int("CFDP_TILE_MAX_FACES", 2688);
  E.sopt.max_local = env_int("CFDP_TILE_MAX_LOCAL", 768);
  E.sopt.order = env_int("CFDP_TILE_ORDER", 0);
  E.sopt.sort_in_tile = env_int("CFDP_SORT_IN_TILE", 1);
  E.sopt.flux_blob = env_int("CFDP_FLUX_BLOB", 1);
  E.sopt.refine_rounds = env_int("CFDP_PLACE_REFINE", 0);
  E.sopt.bank_placement = env_int("CFDP_BANK_PLACEMENT", 1);
  E.sopt.slack_slots = env_int("CFDP_SLACK_SLOTS", 0);
  E.sopt.slack_halo = env_int("CFDP_SLACK_HALO", 0);
  E.sopt.stage_budget = env_int("CFDP_STAGE_BUDGET", 104 * 1024); /* two tiles (two CTAs or two stages) in 228 KB of shared memory per SM */
  E.exact = env_int("CFDP_EXACT", 1);
  E.loopback = nprocs == 1 && env_int("CFDP_LOOPBACK", 0) != 0;
  E.configured = true;
  return 0;
}

extern "C" void cfdp_set_resident(int r) { g_eng.resident = r ? 1 : 0; }
extern "C" void cfdp_set_exact(int x) { g_eng.exact = x ? 1 : 0; }

extern "C" int cfdp_nccl_get_unique_id(void *id128)
{
  nccl_load();
  nccl_uid id;
  NCCL_CHECK(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof id);
  return 0;
}

extern "C" int cfdp_nccl_init(const void *id128)
{
  Engine &E = g_eng;
  ASSERT(E.configured);
  if (E.nprocs == 1 || E.comm) return 0;
  ensure_device();
  nccl_load();
  nccl_uid id; memcpy(&id, id128, sizeof id);
  NCCL_CHECK(g_nccl.CommInitRank(&E.comm, E.nprocs, id, E.proc_rank));
  return 0;
}

/* pure-C drivers (no Python plumbing): rank 0 publishes the NCCL id in a file.  The file lives in a per-user 0700
 * directory, is created with O_EXCL under a temporary name and renamed into place, carries a magic word, the launcher's
 * pid and a timestamp that readers check (a stale file of an earlier run under a recycled pid / port is ignored), and is
 * unlinked by rank 0 once every rank has joined the communicator. */
struct NcclIdFile { unsigned long long magic; long long ppid; long long t_sec; nccl_uid id; };
static void nccl_file_bootstrap(void)
{
  Engine &E = g_eng;
  if (E.nprocs == 1 || E.comm) return;
  char dir[400], path[512];
  const char *f = getenv("CFDP_NCCL_ID_FILE");
  if (f && *f) snprintf(path, sizeof path, "%s", f);
  else {
    const char *base = getenv("XDG_RUNTIME_DIR");
    snprintf(dir, sizeof dir, "%s/cfdp_b200_%d", (base && *base) ? base : "/tmp", (int)getuid());
    if (mkdir(dir, 0700) != 0 && errno != EEXIST) { fprintf(stderr, "Error: cannot create %s [%s:%i]\n", dir, __FILE__, __LINE__); exit(EXIT_FAILURE); }
    struct stat sb;
    ASSERT(stat(dir, &sb) == 0 && S_ISDIR(sb.st_mode) && sb.st_uid == getuid() && (sb.st_mode & 077) == 0); /* ours, private */
    snprintf(path, sizeof path, "%s/nccl_id_%s_%d", dir, getenv("MASTER_PORT") ? getenv("MASTER_PORT") : "29500", (int)getppid());
  }
  const unsigned long long MAGIC = 0x4346445042323030ull; /* "CFDPB200" */
  NcclIdFile rec;
  if (E.proc_rank == 0) {
    memset(&rec, 0, sizeof rec);
    rec.magic = MAGIC; rec.ppid = (long long)getppid(); rec.t_sec = (long long)time(NULL);
    cfdp_nccl_get_unique_id(&rec.id);
    char tmp[600]; snprintf(tmp, sizeof tmp, "%s.tmp.%d", path, (int)getpid());
    unlink(path);                                            /* whatever an earlier run left behind */
    const int fd = open(tmp, O_WRONLY | O_CREAT | O_EXCL, 0600);
    ASSERT(fd >= 0);
    ASSERT(write(fd, &rec, sizeof rec) == (ssize_t)sizeof rec);
    close(fd);
    ASSERT(rename(tmp, path) == 0);
  } else {
    bool ok = false;
    for (int i = 0; i < 6000 && !ok; i++) {
      FILE *fp = fopen(path, "rb");
      if (fp) {
        ok = fread(&rec, sizeof rec, 1, fp) == 1 && rec.magic == MAGIC && rec.ppid == (long long)getppid() &&
             llabs((long long)time(NULL) - rec.t_sec) < 600;  /* this launch, not a leftover */
        fclose(fp);
      }
      if (!ok) usleep(10000);
    }
    ASSERT(ok);
  }
  cfdp_nccl_init(&rec.id);                                   /* collective: returns when every rank has joined */
  if (E.proc_rank == 0) unlink(path);
}

/* setup-time exchange of int messages with other processes (comm_data.c:195-250).  Default
 * transport: NCCL.  The host plumbing may install its own (torch.distributed / gloo on machines
 * without a GPU) with cfdp_set_int_exchange(). */
static cfdp_int_exchange_fn g_int_exchange = nullptr;
extern "C" void cfdp_set_int_exchange(cfdp_int_exchange_fn fn) { g_int_exchange = fn; }

void engine_exchange_ints(const std::vector<int> &peer, const std::vector<const int *> &sbuf, const std::vector<int> &scount,
                          const std::vector<int *> &rbuf, const std::vector<int> &rcount)
{
  Engine &E = g_eng;
  const size_t n = peer.size();
  if (E.nprocs == 1) { /* loopback: this process is its own peer; the i-th send pairs with the i-th receive */
    std::vector<size_t> si, ri;
    for (size_t i = 0; i < n; i++) { ASSERT(peer[i] == E.proc_rank); if (scount[i]) si.push_back(i); if (rcount[i]) ri.push_back(i); }
    ASSERT(si.size() == ri.size());
    for (size_t j = 0; j < si.size(); j++) { ASSERT(scount[si[j]] == rcount[ri[j]]); memcpy(rbuf[ri[j]], sbuf[si[j]], (size_t)scount[si[j]] * sizeof(int)); }
    return;
  }
  if (g_int_exchange) {
    std::vector<const int *> sb(sbuf); std::vector<int *> rb(rbuf);
    g_int_exchange((int)n, peer.data(), sb.data(), scount.data(), rb.data(), rcount.data());
    return;
  }
  ensure_device();
  if (!E.comm) nccl_file_bootstrap();
  ASSERT(E.comm != nullptr);
  std::vector<int *> dbuf(n, nullptr);
  for (size_t i = 0; i < n; i++) {
    const int cnt = scount[i] + rcount[i];
    CUDA_CHECK(cudaMalloc(&dbuf[i], (size_t)(cnt > 0 ? cnt : 1) * sizeof(int)));
    if (scount[i]) CUDA_CHECK(cudaMemcpyAsync(dbuf[i], sbuf[i], (size_t)scount[i] * sizeof(int), cudaMemcpyHostToDevice, E.s_comm));
  }
  NCCL_CHECK(g_nccl.GroupStart());
  for (size_t i = 0; i < n; i++) {
    if (scount[i]) NCCL_CHECK(g_nccl.Send(dbuf[i], (size_t)scount[i], NCCL_INT32, peer[i], E.comm, E.s_comm));
    if (rcount[i]) NCCL_CHECK(g_nccl.Recv(dbuf[i], (size_t)rcount[i], NCCL_INT32, peer[i], E.comm, E.s_comm));
  }
  NCCL_CHECK(g_nccl.GroupEnd());
  for (size_t i = 0; i < n; i++)
    if (rcount[i]) CUDA_CHECK(cudaMemcpyAsync(rbuf[i], dbuf[i], (size_t)rcount[i] * sizeof(int), cudaMemcpyDeviceToHost, E.s_comm));
  CUDA_CHECK(cudaStreamSynchronize(E.s_comm));
  for (size_t i = 0; i < n; i++) CUDA_CHECK(cudaFree(dbuf[i]));
}

/* ------------------------------------------------------------------------------------------
 * init_threads: the reference builds its CPU schedule here (threads.c:730-788).  We register the
 * domain; the GPU schedule of all hosted domains is built together in cfdp_commit().
 * NTHREADS is accepted for signature compatibility and ignored.
 * ---------------------------------------------------------------------------------------- */
extern "C" void init_threads(comm_data *cd, solver_data *sd, int NTHREADS)
{
  (void)NTHREADS;
  ASSERT(cd != NULL);
  ASSERT(sd != NULL);
  Domain *d = engine_find_domain(cd);
  ASSERT(d != NULL);
  ASSERT(!g_eng.planned);
  ASSERT(sd->nownpoints == cd->nownpoints);
  if (cd->ndomains > 1) ASSERT(d->tables_done); /* compute_communication_tables first */
  d->sd = sd;
  d->threads_inited = true;
}

/* ------------------------------------------------------------------------------------------
 * Kernels
 * ---------------------------------------------------------------------------------------- */
/* dst[dst_rows ? dst_rows[i] : i][:] = src[src_rows ? src_rows[i] : i][:]  -- pack / unpack /
 * on-device halo copy (threads.c:791-869: raw copies of dim2 = 21 doubles per point) */
__global__ void rows_copy_kernel(double *__restrict__ dst, const uint32_t *__restrict__ dst_rows,
                                 const double *__restrict__ src, const uint32_t *__restrict__ src_rows, long long nrows, int width,
                                 double scale)
{
  const long long total = nrows * width;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / width; const int c = (int)(i - r * width);
    const size_t d = (dst_rows ? (size_t)dst_rows[r] : (size_t)r) * width + c;
    const size_t s = (src_rows ? (size_t)src_rows[r] : (size_t)r) * width + c;
    dst[d] = scale == 1.0 ? src[s] : __dmul_rn(src[s], scale);
  }
}

static void launch_rows_copy(double *dst, const uint32_t *dst_rows, const double *src, const uint32_t *src_rows,
                             long long nrows, int width, cudaStream_t st, double scale = 1.0)
{
  if (nrows <= 0) return;
  /* 64-thread blocks: small enough (registers) to become resident next to two gradient CTAs, so that pack / unpack
   * of the overlapped exchange do not have to wait for a gradient CTA to retire */
  const long long total = nrows * width;
  long long blocks = (total + 63) / 64;
  static const long long cap = env_int("CFDP_COPY_BLOCKS", 148 * 32);
  if (blocks > cap) blocks = cap;
  rows_copy_kernel<<<(unsigned)blocks, 64, 0, st>>>(dst, dst_rows, src, src_rows, nrows, width, scale);
  CUDA_CHECK(cudaGetLastError());
  g_eng.launches++;
}

/* refresh the packed halo rows of tiles [tile0, tile0 + ntiles) from the device var rows (after every upload of var) */
static void launch_halo_pack(long long tile0, long long ntiles, cudaStream_t st)
{
  Engine &E = g_eng;
  if (ntiles <= 0) return;
  const unsigned grid = (unsigned)std::min<long long>(ntiles, 148 * 16);
  ggk::halo_pack_kernel<<<grid, 256, 0, st>>>(E.d_tiles + tile0, (int)ntiles, E.d_blob, E.d_var, E.d_hhalo);
  CUDA_CHECK(cudaGetLastError());
  E.launches++;
}
static void launch_halo_pack_domain(const Domain *d, cudaStream_t st)
{
  launch_halo_pack(d->tile0_b, d->sch.nboundary, st);
  launch_halo_pack(d->tile0_i, d->sch.ntiles - d->sch.nboundary, st);
}

static void launch_gradient(long long tile0, long long ntiles, cudaStream_t st, int nsignal = 0, bool exports = false, bool direct = false)
{
  Engine &E = g_eng;
  if (ntiles <= 0) return;
  if (E.kernel_version == 2) {
    ggk::PipeLayout &P = E.pipe;
    P.nsignal = nsignal; P.progress = E.d_progress;
    P.tile_base = (int)tile0;
    P.nexport = (exports && E.fused_pack) ? (int)E.nbtiles : 0;
    P.exp_off = E.d_exp_off; P.exp_src = direct ? E.d_exp_src_direct : E.d_exp_src; P.exp_dst = direct ? E.d_exp_dst_direct : E.d_exp_dst;
    P.exp_base[0] = E.d_grad; P.exp_base[1] = E.d_sendbuf;
    P.sig_off = direct ? E.d_sig_off : nullptr; P.sig_ent = E.d_sig_ent;
    unsigned grid;
    if (E.persistent > 0) { /* interleaved: CTA b walks tiles b, b + grid, ...: the tiles in flight at any time are neighbours in the tile order */
      grid = (unsigned)std::min<long long>(ntiles, (long long)E.persistent);
      P.cstride = 1; P.istride = (int)grid; P.maxcount = 0x7FFFFFFF;
    } else {
      grid = (unsigned)((ntiles + E.chunk - 1) / E.chunk);
      P.cstride = E.chunk; P.istride = 1; P.maxcount = E.chunk;
    }
    /* direct halo stores: boundary tiles dealt out over the first CFDP_DIRECT_SPREAD percent of the walk (default 0 = boundary tiles
     * first; spreading measured SLOWER at 2 GPUs: every CTA then pays one system-scope release, profiles/README.md) */
    P.spread_m = 0; P.spread_nb = 0;
    if (direct && tile0 == 0 && ntiles == E.ntiles && E.nbtiles > 0) {
      const long long pct = env_int("CFDP_DIRECT_SPREAD", 0);
      const long long m = E.ntiles * pct / 100 / E.nbtiles;
      if (m > 1) { P.spread_m = (int)m; P.spread_nb = (int)E.nbtiles; }
    }
    P.variant = env_int("CFDP_VARIANT", 1); /* 1 = L2 prefetch of the next tile's late-fetched blob head (+5.5 .. 7 %, profiles/README.md) */
#define CFDP_LAUNCH_PIPE(EX, NC) ggk::gg_tile_pipe_kernel<EX, NC><<<grid, E.block_threads, E.smem_bytes, st>>>(E.d_tiles + tile0, (int)ntiles, E.d_blob, E.d_var, E.d_hhalo, E.d_pvol, E.d_grad, P)
    if (E.ctas_per_sm == 4) { if (E.exact) CFDP_LAUNCH_PIPE(true, 4); else CFDP_LAUNCH_PIPE(false, 4); }
    else if (E.ctas_per_sm == 3) { if (E.exact) CFDP_LAUNCH_PIPE(true, 3); else CFDP_LAUNCH_PIPE(false, 3); }
    else { if (E.exact) CFDP_LAUNCH_PIPE(true, 2); else CFDP_LAUNCH_PIPE(false, 2); }
#undef CFDP_LAUNCH_PIPE
  } else {
    if (E.exact)
      ggk::gg_tile_kernel<true><<<(unsigned)ntiles, E.block_threads, E.smem_v1, st>>>(E.d_tiles + tile0, E.d_blob, E.d_var, E.d_pvol, E.d_grad, E.region0_doubles);
    else
      ggk::gg_tile_kernel<false><<<(unsigned)ntiles, E.block_threads, E.smem_v1, st>>>(E.d_tiles + tile0, E.d_blob, E.d_var, E.d_pvol, E.d_grad, E.region0_doubles);
  }
  CUDA_CHECK(cudaGetLastError());
  E.launches++;
}

/* pseudo flux of tiles [tile0, tile0 + ntiles) from the device grad rows (flux.c:111-201) */
static void flux_prepare(void)
{
  Engine &E = g_eng;
  if (E.d_flux) return;
  ASSERT(E.committed);
  ASSERT((int)E.flux_smem <= 227 * 1024 - 64);
  CUDA_CHECK(cudaMalloc(&E.d_flux, (size_t)E.rows * NFLUX * sizeof(double)));
  CUDA_CHECK(cudaMemset(E.d_flux, 0, (size_t)E.rows * NFLUX * sizeof(double)));
  CUDA_CHECK(cudaFuncSetAttribute(ggk::psd_flux_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)E.flux_smem));
  CUDA_CHECK(cudaFuncSetAttribute(ggk::psd_flux_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)E.flux_smem));
  CUDA_CHECK(cudaFuncSetAttribute(ggk::psd_flux_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)E.flux_smem));
  CUDA_CHECK(cudaFuncSetAttribute(ggk::psd_flux_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)E.flux_smem));
}
static void launch_flux(long long tile0, long long ntiles, cudaStream_t st)
{
  Engine &E = g_eng;
  if (ntiles <= 0) return;
  flux_prepare();
  const int version = env_int("CFDP_FLUX_KERNEL", 2);
  const TileDesc *tiles = E.d_ftiles ? E.d_ftiles : E.d_tiles;
  const unsigned char *blob = E.d_ftiles ? E.d_fblob : E.d_blob;
  if (version == 2 && E.max_nhalo <= CFDP_FLUX_HALO_PER_THREAD * E.block_threads) {
    const int chunk = std::min(E.chunk, CFDP_FLUX_MAX_CHUNK);
    const unsigned grid = (unsigned)((ntiles + chunk - 1) / chunk);
    const int fvariant = env_int("CFDP_FLUX_VARIANT", 2); /* 2 = the next tile's grad rows are asked into L2 before this tile's walk (+3 %) */
    if (E.exact)
      ggk::psd_flux_pipe_kernel<true><<<grid, E.block_threads, E.flux_smem, st>>>(tiles + tile0, (int)ntiles, chunk, blob, E.d_grad, E.d_flux, fvariant);
    else
      ggk::psd_flux_pipe_kernel<false><<<grid, E.block_threads, E.flux_smem, st>>>(tiles + tile0, (int)ntiles, chunk, blob, E.d_grad, E.d_flux, fvariant);
  } else if (E.exact)
    ggk::psd_flux_tile_kernel<true><<<(unsigned)ntiles, E.block_threads, E.flux_smem, st>>>(tiles + tile0, blob, E.d_grad, E.d_flux);
  else
    ggk::psd_flux_tile_kernel<false><<<(unsigned)ntiles, E.block_threads, E.flux_smem, st>>>(tiles + tile0, blob, E.d_grad, E.d_flux);
  CUDA_CHECK(cudaGetLastError());
  E.launches++;
}

/* ------------------------------------------------------------------------------------------
 * commit: schedules, device layout, exchange plan
 * ---------------------------------------------------------------------------------------- */
template <typename T>
static T *upload(const std::vector<T> &v)
{
  T *d = nullptr;
  CUDA_CHECK(cudaMalloc(&d, (v.size() ? v.size() : 1) * sizeof(T)));
  if (!v.empty()) CUDA_CHECK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

/* host part: schedules, unified rows, tile list, exchange row lists.  No device needed. */
extern "C" void cfdp_plan(void)
{
  Engine &E = g_eng;
  if (E.planned) return;
  ASSERT(E.configured);
  ASSERT((int)E.doms.size() == E.per_proc);
  for (Domain *d : E.doms) { ASSERT(d->threads_inited && d->sd != NULL); }
  const int nh = (int)E.doms.size();

  /* 1. face schedules: domains side by side when several are hosted (the cores that are left over work inside each
   * domain's builder, nested regions); very large domains one or two at a time: the builder needs ~0.5 KB per point */
  const bool lean = env_int("CFDP_LEAN_HOST", 0) != 0;
  long long total_pts = 0;
  for (Domain *d : E.doms) total_pts += d->sd->nallpoints;
  auto build_one = [&](int i) {
    Domain *d = E.doms[(size_t)i];
    build_schedule(d->sd, d->cd, E.sopt, d->sch);
    if (lean) { /* the schedule holds everything the device needs: release the mesh arrays (CFDP_LEAN_HOST) */
      free(d->sd->fpoint); free(d->sd->fnormal); d->sd->fpoint = nullptr; d->sd->fnormal = nullptr;
      std::vector<int>().swap(d->sch.tile_face_ids); std::vector<int>().swap(d->sch.tile_halo_pts); /* introspection lists (cfdp_get_tile) */
    }
  };
  if (nh > 1) {
    const int ncores = omp_get_max_threads();
    const int outer = std::min(total_pts > 80000000LL ? 2 : nh, std::min(nh, ncores)), inner = std::max(1, ncores / outer);
    const int levels = omp_get_max_active_levels();
    if (inner > 1) omp_set_max_active_levels(std::max(levels, 2));
#pragma omp parallel for schedule(dynamic, 1) num_threads(outer)
    for (int i = 0; i < nh; i++) {
      omp_set_num_threads(inner); /* this thread's nested regions */
      build_one(i);
    }
    omp_set_max_active_levels(levels);
  } else {
    build_one(0);
  }

  /* 2. unified rows and tile list: boundary tiles of all domains first */
  E.rows = 0; E.ntiles = 0; E.nbtiles = 0; E.blob_bytes = 0;
  for (Domain *d : E.doms) {
    d->rowbase = E.rows; E.rows += d->sch.nrows;
    E.ntiles += d->sch.ntiles; E.nbtiles += d->sch.nboundary;
    E.max_nfaces = std::max(E.max_nfaces, d->sch.max_nfaces); E.max_nloc = std::max(E.max_nloc, d->sch.max_nloc);
    for (int n : d->sch.tile_npts) E.max_npts = std::max(E.max_npts, n);
    E.max_stage = std::max(E.max_stage, (size_t)d->sch.nall * (lean ? NGRAD : CFDP_DIM2) * sizeof(double));
    E.max_blob = std::max(E.max_blob, d->sch.max_blob);
    for (int n : d->sch.tile_nhpos) E.max_nhalo = std::max(E.max_nhalo, n);
    E.nfaces += d->sch.nfaces_computed; E.nown += d->sch.nown; E.nall += d->sch.nall;
    E.tile_faces += d->sch.tile_faces; E.halo_refs += d->sch.halo_refs;
    E.alg_bytes += d->sch.nfaces_computed * 32 + (long long)d->sch.nall * 56 + (long long)d->sch.nown * 176;
    /* pseudo flux: fpoint + fnormal per face, 72 of the 168 bytes of every grad row read once, 24 bytes of psd_flux written per own point */
    E.flux_alg_bytes += d->sch.nfaces_computed * 32 + (long long)d->sch.nall * 72 + (long long)d->sch.nown * 24;
  }
  ASSERT(E.rows < 0x7FFFFFF0LL);
  E.h_tiles.resize((size_t)E.ntiles);
  E.blob_base.resize((size_t)nh);
  for (int i = 0; i < nh; i++) { E.blob_base[i] = E.blob_bytes; E.blob_bytes += E.doms[i]->sch.blob.size(); }
  const bool fb = E.sopt.flux_blob != 0;
  E.h_ftiles.assign(fb ? (size_t)E.ntiles : 0, TileDesc{}); E.fblob_base.assign((size_t)nh, 0); E.fblob_bytes = 0;
  for (int i = 0; i < nh; i++) { E.fblob_base[i] = E.fblob_bytes; E.fblob_bytes += E.doms[i]->sch.fblob.size(); }
  long long tb = 0, ti = E.nbtiles;
  E.point_of_row.resize((size_t)nh);
  for (int i = 0; i < nh; i++) {
    Domain *d = E.doms[i]; DomainSchedule &s = d->sch;
    d->tile0_b = tb; d->tile0_i = ti;
    for (int k = 0; k < s.ntiles; k++) {
      TileDesc t;
      t.row0 = (uint32_t)(d->rowbase + s.tile_row0[k]);
      t.npts = (uint16_t)s.tile_npts[k]; t.nhalo = (uint16_t)s.tile_nhpos[k];     /* positions, gaps hold 0xFFFFFFFF */
      t.nfaces = (uint16_t)s.tile_nslots[k]; t.zslot = (uint16_t)s.tile_zslot[k]; t.maxdeg = (uint16_t)s.tile_maxdeg[k];
      ASSERT((E.blob_base[i] + s.tile_blob[k]) % 128 == 0);
      t.blob128 = (uint32_t)((E.blob_base[i] + s.tile_blob[k]) / 128); t.hrow0 = 0; /* assigned below, in launch order */
      t.blob_bytes = (uint32_t)(s.tile_blob[(size_t)k + 1] - s.tile_blob[k]);
      t.npad = (uint16_t)align_up((size_t)s.tile_npts[k], 32);
      t.halo_off = (uint32_t)blob_halo_off(t.nfaces);
      /* rebase the tile's halo rows from domain-relative to device rows */
      uint32_t *hr = (uint32_t *)(&s.blob[s.tile_blob[k]] + t.halo_off);
      for (int j = 0; j < s.tile_nhpos[k]; j++) if (hr[j] != 0xFFFFFFFFu) hr[j] += (uint32_t)d->rowbase;
      E.max_footprint = std::max(E.max_footprint, ggk::tile_footprint(t.blob_bytes, t.npts, t.nhalo));
      E.max_hvpv = std::max(E.max_hvpv, ggk::stage_hvar_bytes(t.npts, t.nhalo) + ggk::stage_pvol_bytes(t.npts));
      const size_t slot = (size_t)(k < s.nboundary ? tb++ : ti++);
      E.h_tiles[slot] = t;
      if (fb) { /* the tile's pseudo-flux blob: same format, fewer faces / halo rows / adjacency rows */
        TileDesc f = t;
        f.nhalo = (uint16_t)s.ftile_nhalo[k]; f.nfaces = (uint16_t)s.ftile_nfaces[k]; f.zslot = 0; f.maxdeg = (uint16_t)s.ftile_maxdeg[k];
        ASSERT((E.fblob_base[i] + s.ftile_blob[k]) % 128 == 0);
        f.blob128 = (uint32_t)((E.fblob_base[i] + s.ftile_blob[k]) / 128); f.hrow0 = 0;
        f.blob_bytes = (uint32_t)(s.ftile_blob[(size_t)k + 1] - s.ftile_blob[k]);
        f.halo_off = (uint32_t)blob_halo_off(f.nfaces);
        uint32_t *fh = (uint32_t *)(&s.fblob[s.ftile_blob[k]] + f.halo_off);
        for (int j = 0; j < s.ftile_nhalo[k]; j++) if (fh[j] != 0xFFFFFFFFu) fh[j] += (uint32_t)d->rowbase;
        E.h_ftiles[slot] = f;
        E.flux_smem = std::max(E.flux_smem, ggk::flux_footprint(f.blob_bytes, f.halo_off, f.npts, f.nhalo));
      } else {
        E.flux_smem = std::max(E.flux_smem, ggk::flux_footprint(t.blob_bytes, t.halo_off, t.npts, t.nhalo));
      }
    }
    E.point_of_row[i].assign((size_t)s.nrows, -1);
    for (int p = 0; p < s.nall; p++) E.point_of_row[i][(size_t)s.row_of_point[p]] = p;
  }
  ASSERT(tb == E.nbtiles && ti == E.ntiles);
  E.halo_rows = 0;
  for (long long t = 0; t < E.ntiles; t++) { ASSERT(E.halo_rows < 0xFFFFFFF0LL); E.h_tiles[(size_t)t].hrow0 = (uint32_t)E.halo_rows; E.halo_rows += E.h_tiles[(size_t)t].nhalo; }

  /* 3. exchange plan (thread_comm.c:27-432 / threads.c:571-726 flattened into device row lists) */
  struct Seg { int src, dst; const std::vector<uint32_t> *rows; };
  std::map<int, std::vector<Seg>> send_segs, recv_segs; /* by peer process */
  for (int i = 0; i < nh; i++) {
    Domain *d = E.doms[i]; comm_data *a = d->cd;
    if (a->ndomains == 1) continue;
    for (int sl = 0; sl < a->ncommdomains; sl++) {
      const int k = a->commpartner[sl];
      std::vector<uint32_t> &sr = E.send_rows_of[{a->iProc, k}], &rr = E.recv_rows_of[{a->iProc, k}];
      sr.resize((size_t)a->sendcount[k]); rr.resize((size_t)a->recvcount[k]);
      for (int j = 0; j < a->sendcount[k]; j++) sr[j] = (uint32_t)(d->rowbase + d->sch.row_of_point[a->sendindex[k][j]]);
      for (int j = 0; j < a->recvcount[k]; j++) rr[j] = (uint32_t)(d->rowbase + d->sch.row_of_point[a->recvindex[k][j]]);
    }
  }
  for (int i = 0; i < nh; i++) {
    Domain *d = E.doms[i]; comm_data *a = d->cd;
    if (a->ndomains == 1) continue;
    for (int sl = 0; sl < a->ncommdomains; sl++) {
      const int k = a->commpartner[sl];
      if (engine_domain_by_id(k) && !E.loopback) { /* ghost rows of a owned by k, both on this GPU: one device copy */
        const std::vector<uint32_t> &dst = E.recv_rows_of[{a->iProc, k}], &src = E.send_rows_of[{k, a->iProc}];
        ASSERT(dst.size() == src.size());
        E.h_loc_dst.insert(E.h_loc_dst.end(), dst.begin(), dst.end()); E.h_loc_src.insert(E.h_loc_src.end(), src.begin(), src.end());
      } else {
        const int q = engine_proc_of_domain(k); /* loopback: q is this process */
        if (a->sendcount[k]) send_segs[q].push_back({a->iProc, k, &E.send_rows_of[{a->iProc, k}]});
        if (a->recvcount[k]) recv_segs[q].push_back({k, a->iProc, &E.recv_rows_of[{a->iProc, k}]});
      }
    }
  }
  E.n_local = (long long)E.h_loc_dst.size();
  std::vector<int> procs;
  for (auto &kv : send_segs) procs.push_back(kv.first);
  for (auto &kv : recv_segs) if (std::find(procs.begin(), procs.end(), kv.first) == procs.end()) procs.push_back(kv.first);
  std::sort(procs.begin(), procs.end());
  auto by_key = [](const Seg &x, const Seg &y) { return x.src != y.src ? x.src < y.src : x.dst < y.dst; };
  for (int q : procs) {
    PeerPlan pp; pp.proc = q;
    pp.send_off = (long long)E.h_send_rows.size(); pp.recv_off = (long long)E.h_recv_rows.size();
    std::vector<Seg> &ss = send_segs[q], &rs = recv_segs[q];
    std::sort(ss.begin(), ss.end(), by_key); std::sort(rs.begin(), rs.end(), by_key); /* same order on both sides */
    for (const Seg &sg : ss) E.h_send_rows.insert(E.h_send_rows.end(), sg.rows->begin(), sg.rows->end());
    for (const Seg &sg : rs) E.h_recv_rows.insert(E.h_recv_rows.end(), sg.rows->begin(), sg.rows->end());
    pp.send_rows = (long long)E.h_send_rows.size() - pp.send_off; pp.recv_rows = (long long)E.h_recv_rows.size() - pp.recv_off;
    /* halo relations are symmetric (a face joins a point of each side): the stage-parity double buffering of the put + notify
     * window relies on it -- a peer that only received could be overrun by a sender two stages ahead */
    ASSERT((pp.send_rows > 0) == (pp.recv_rows > 0));
    E.peers.push_back(pp);
  }
  E.n_send = (long long)E.h_send_rows.size(); E.n_recv = (long long)E.h_recv_rows.size();
  ASSERT(E.peers.size() <= (size_t)MAX_PEERS);

  /* 4. export lists (fused pack, threads.c:187-249 "pack while computing"): for every boundary tile the rows of it
   * that some other domain needs, with their destination: a ghost row of a domain hosted on this GPU (bit 31
   * clear, row of grad) or a slot of the packed send buffer (bit 31 set) */
  {
    std::vector<std::pair<uint32_t, int>> by_row((size_t)E.ntiles);
    for (long long t = 0; t < E.ntiles; t++) by_row[(size_t)t] = {E.h_tiles[(size_t)t].row0, (int)t};
    std::sort(by_row.begin(), by_row.end());
    auto tile_of = [&](uint32_t row) {
      auto it = std::upper_bound(by_row.begin(), by_row.end(), std::make_pair(row, 0x7FFFFFFF));
      ASSERT(it != by_row.begin());
      --it;
      const TileDesc &td = E.h_tiles[(size_t)it->second];
      ASSERT(row >= td.row0 && row < td.row0 + td.npts);
      return it->second;
    };
    const size_t nexp = E.h_loc_src.size() + E.h_send_rows.size();
    std::vector<int> etile(nexp); std::vector<uint32_t> esrc(nexp), edst(nexp);
    size_t n = 0;
    std::vector<unsigned char> ekind(nexp);
    for (size_t i = 0; i < E.h_loc_src.size(); i++, n++) { etile[n] = tile_of(E.h_loc_src[i]); esrc[n] = E.h_loc_src[i]; edst[n] = E.h_loc_dst[i]; ekind[n] = 0; }
    for (size_t i = 0; i < E.h_send_rows.size(); i++, n++) { etile[n] = tile_of(E.h_send_rows[i]); esrc[n] = E.h_send_rows[i]; edst[n] = (uint32_t)i; ekind[n] = 1; }
    E.h_exp_off.assign((size_t)E.nbtiles + 1, 0);
    for (size_t i = 0; i < nexp; i++) { ASSERT(etile[i] < E.nbtiles); E.h_exp_off[(size_t)etile[i] + 1]++; } /* send points live in boundary tiles */
    for (long long t = 0; t < E.nbtiles; t++) E.h_exp_off[(size_t)t + 1] += E.h_exp_off[(size_t)t];
    E.h_exp_src.resize(nexp); E.h_exp_dst.resize(nexp);
    std::vector<uint32_t> cur(E.h_exp_off.begin(), E.h_exp_off.end() - 1);
    for (size_t i = 0; i < nexp; i++) {
      const uint32_t pos = cur[(size_t)etile[i]]++;
      E.h_exp_src[pos] = (esrc[i] - E.h_tiles[(size_t)etile[i]].row0) | ((uint32_t)ekind[i] << 16);   /* tile-local point | destination array << 16 */
      E.h_exp_dst[pos] = edst[i];
    }
  }
  E.planned = true;
}

/* which gradient kernel runs and how its grid walks the tiles.  version 2 = production (gg_tile_pipe_kernel),
 * 1 = one tile per CTA (second, independent implementation for cross-checks);
 * chunk = consecutive tiles per CTA; persistent > 0 = that many CTAs walk all tiles, interleaved */
static void configure_kernel(int version, int chunk, int persistent)
{
  Engine &E = g_eng;
  /* pipelined kernels: one stage [blob | var rows | volumes] per CTA, two CTAs per SM.  The production kernel keeps the
   * var rows at the END of the stage: they must never overlap the rows the previous tile is still storing (one zone per warp) */
  E.pipe.stage_bytes = std::max(E.max_footprint, (uint32_t)((E.block_threads / 32) * CFDP_ZONE_BYTES) + E.max_hvpv);
  E.kernel_version = version;
  E.chunk = std::max(1, chunk);
  E.persistent = std::max(0, persistent);
  const int smem_limit = 227 * 1024 - 256;
  if (E.kernel_version == 2 && (int)E.pipe.stage_bytes > smem_limit) {
    fprintf(stderr, "cfdp: pipelined kernel needs %u B of shared memory per stage, falling back to the one-tile-per-CTA kernel "
                    "(lower CFDP_TILE_POINTS to avoid this)\n", E.pipe.stage_bytes);
    E.kernel_version = 1;
  }
  ASSERT(E.smem_v1 <= smem_limit);
  E.smem_bytes = E.kernel_version == 2 ? (int)E.pipe.stage_bytes : E.smem_v1;
  CUDA_CHECK(cudaFuncSetAttribute(ggk::gg_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, E.smem_v1));
  CUDA_CHECK(cudaFuncSetAttribute(ggk::gg_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, E.smem_v1));
  if (E.kernel_version == 2) {
    /* CFDP_CTAS = 3 / 4: three CTAs of <= 160 threads (15 warps, <= 75 KB each) or four of <= 128 (16 warps, <= 56 KB) per SM
     * instead of two of 256; the register file holds 128 registers per thread in all three shapes */
    const int want = env_int("CFDP_CTAS", 2);
    E.ctas_per_sm = (want == 4 && E.block_threads <= 128 && E.smem_bytes <= 56 * 1024) ? 4 : (want == 3 && E.block_threads <= 160 && E.smem_bytes <= 75 * 1024) ? 3 : 2;
    CUDA_CHECK(cudaFuncSetAttribute(ggk::gg_tile_pipe_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, E.smem_bytes));
    CUDA_CHECK(cudaFuncSetAttribute(ggk::gg_tile_pipe_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, E.smem_bytes));
    if (E.ctas_per_sm == 3) {
      CUDA_CHECK(cudaFuncSetAttribute(ggk::gg_tile_pipe_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, E.smem_bytes));
      CUDA_CHECK(cudaFuncSetAttribute(ggk::gg_tile_pipe_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, E.smem_bytes));
    } else if (E.ctas_per_sm == 4) {
      CUDA_CHECK(cudaFuncSetAttribute(ggk::gg_tile_pipe_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, E.smem_bytes));
      CUDA_CHECK(cudaFuncSetAttribute(ggk::gg_tile_pipe_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, E.smem_bytes));
    }
  }
}

extern "C" int cfdp_set_kernel(int version, int chunk, int persistent)
{
  Engine &E = g_eng;
  if (!E.committed || version < 1 || version > 2) return -1;
  cfdp_device_synchronize();
  configure_kernel(version, chunk, persistent);
  return E.kernel_version;
}

static void ipc_setup(void);
typedef CUresult (*wait_value64_fn)(CUstream, CUdeviceptr, cuuint64_t, unsigned int);
static wait_value64_fn get_wait_value64(void);

/* device part: allocate, upload, configure the kernel */
extern "C" void cfdp_commit(void)
{
  Engine &E = g_eng;
  if (E.committed) return;
  cfdp_plan();
  ensure_device();
  const int nh = (int)E.doms.size();
  CUDA_CHECK(cudaMalloc(&E.d_blob, E.blob_bytes ? E.blob_bytes : 16));
  for (int i = 0; i < nh; i++) {
    DomainSchedule &s = E.doms[i]->sch;
    CUDA_CHECK(cudaMemcpy(E.d_blob + E.blob_base[i], s.blob.data(), s.blob.size(), cudaMemcpyHostToDevice));
    std::vector<unsigned char>().swap(s.blob); /* the host copy is not needed any more */
  }
  E.d_tiles = upload(E.h_tiles);
  if (E.sopt.flux_blob) {
    CUDA_CHECK(cudaMalloc(&E.d_fblob, E.fblob_bytes ? E.fblob_bytes : 16));
    for (int i = 0; i < nh; i++) {
      DomainSchedule &s = E.doms[i]->sch;
      CUDA_CHECK(cudaMemcpy(E.d_fblob + E.fblob_base[i], s.fblob.data(), s.fblob.size(), cudaMemcpyHostToDevice));
      std::vector<unsigned char>().swap(s.fblob);
    }
    E.d_ftiles = upload(E.h_ftiles);
  }

  CUDA_CHECK(cudaMalloc(&E.d_var, (size_t)E.rows * NGRAD * sizeof(double)));
  CUDA_CHECK(cudaMalloc(&E.d_grad, (size_t)E.rows * CFDP_DIM2 * sizeof(double)));
  CUDA_CHECK(cudaMalloc(&E.d_pvol, (size_t)E.rows * sizeof(double)));
  CUDA_CHECK(cudaMalloc(&E.d_hhalo, (size_t)std::max<long long>(E.halo_rows, 2) * NGRAD * sizeof(double)));
  CUDA_CHECK(cudaMemset(E.d_var, 0, (size_t)E.rows * NGRAD * sizeof(double)));
  CUDA_CHECK(cudaMemset(E.d_grad, 0, (size_t)E.rows * CFDP_DIM2 * sizeof(double)));
  E.stage_bytes = E.max_stage;
  CUDA_CHECK(cudaMalloc(&E.d_stage, E.stage_bytes));
  E.d_rowmap.resize((size_t)nh);
  {
    std::vector<double> pv((size_t)E.rows, 1.0); /* padding rows: volume 1 */
    for (int i = 0; i < nh; i++) {
      Domain *d = E.doms[i]; DomainSchedule &s = d->sch;
      std::vector<int> rm((size_t)s.nall);
      for (int p = 0; p < s.nall; p++) { rm[p] = (int)(d->rowbase + s.row_of_point[p]); pv[(size_t)rm[p]] = d->sd->pvolume[p]; }
      E.d_rowmap[i] = upload(rm);
    }
    CUDA_CHECK(cudaMemcpy(E.d_pvol, pv.data(), pv.size() * sizeof(double), cudaMemcpyHostToDevice));
  }

  E.block_threads = (int)align_up((size_t)E.max_npts, 32);
  ASSERT(E.block_threads <= CFDP_MAX_TILE_POINTS);
  ASSERT(E.max_nhalo <= CFDP_MAX_HALO_POS);
  /* v1 (one tile per CTA) */
  E.region0_doubles = (int)align_up((size_t)std::max(E.max_nfaces * 3, E.max_npts * CFDP_DIM2), 2);
  E.smem_v1 = (int)((size_t)E.region0_doubles * 8 + align_up((size_t)E.max_nloc * NGRAD * 8, 16));
  CUDA_CHECK(cudaMalloc(&E.d_progress, 64)); CUDA_CHECK(cudaMemset(E.d_progress, 0, 64)); E.progress_target = 0;
  if (env_int("CFDP_PHASE_PROF", 0)) { CUDA_CHECK(cudaMalloc(&E.pipe.prof, 8 * sizeof(unsigned long long))); CUDA_CHECK(cudaMemset(E.pipe.prof, 0, 8 * sizeof(unsigned long long))); }
  E.fused_signal = env_int("CFDP_FUSED_SIGNAL", 1);
  configure_kernel(env_int("CFDP_KERNEL", 2), env_int("CFDP_CHUNK", 8), env_int("CFDP_PERSISTENT", 0));

  E.d_loc_dst = upload(E.h_loc_dst); E.d_loc_src = upload(E.h_loc_src);
  E.d_exp_off = upload(E.h_exp_off); E.d_exp_src = upload(E.h_exp_src); E.d_exp_dst = upload(E.h_exp_dst);
  E.fused_pack = env_int("CFDP_FUSED_PACK", 1);
  E.d_send_rows = upload(E.h_send_rows); E.d_recv_rows = upload(E.h_recv_rows);
  CUDA_CHECK(cudaMalloc(&E.d_sendbuf, (size_t)std::max<long long>(E.n_send, 1) * CFDP_DIM2 * sizeof(double)));
  CUDA_CHECK(cudaMalloc(&E.d_recvbuf, (size_t)std::max<long long>(E.n_recv, 1) * CFDP_DIM2 * sizeof(double)));
  if (!E.peers.empty() && !E.loopback && !E.comm) {
    if (g_int_exchange) { /* the host plumbing carried the setup handshake itself (cfdp_set_int_exchange) but never called cfdp_nccl_init */
      fprintf(stderr, "Error: %d peer GPU(s) but no NCCL communicator: call cfdp_nccl_init() before cfdp_commit() [%s:%i]\n", (int)E.peers.size(), __FILE__, __LINE__);
      exit(EXIT_FAILURE);
    }
    nccl_file_bootstrap();
  }
  if (!E.peers.empty() && env_int("CFDP_IPC", 1)) ipc_setup();
  CUDA_CHECK(cudaDeviceSynchronize());
  E.committed = true;
  /* var as it stands in the host containers */
  for (Domain *d : E.doms) cfdp_var_to_device(d->sd);
  CUDA_CHECK(cudaStreamSynchronize(E.s_comp));
}

/* ------------------------------------------------------------------------------------------
 * host <-> device mirrors
 * ---------------------------------------------------------------------------------------- */
static int hosted_index(const Domain *d) { for (size_t i = 0; i < g_eng.doms.size(); i++) if (g_eng.doms[i] == d) return (int)i; return -1; }

extern "C" void cfdp_var_to_device(solver_data *sd)
{
  Engine &E = g_eng;
  if (!E.committed) { cfdp_commit(); return; }
  Domain *d = engine_find_domain(sd);
  ASSERT(d != NULL);
  const int i = hosted_index(d);
  const size_t n = (size_t)sd->nallpoints;
  CUDA_CHECK(cudaMemcpyAsync(E.d_stage, &sd->var[0][0], n * NGRAD * sizeof(double), cudaMemcpyHostToDevice, E.s_comp));
  /* the device keeps hvar = 0.5*var in tile order (exact: power-of-two scaling), see gg_kernels.cuh */
  launch_rows_copy(E.d_var, (const uint32_t *)E.d_rowmap[i], E.d_stage, nullptr, (long long)n, NGRAD, E.s_comp, 0.5);
  launch_halo_pack_domain(d, E.s_comp);
}

extern "C" void cfdp_grad_to_host(solver_data *sd)
{
  Engine &E = g_eng;
  ASSERT(E.committed);
  Domain *d = engine_find_domain(sd);
  ASSERT(d != NULL);
  ASSERT(sd->grad != NULL); /* CFDP_LEAN_HOST: there is no host mirror of grad */
  const int i = hosted_index(d);
  const size_t n = (size_t)sd->nallpoints;
  launch_rows_copy(E.d_stage, nullptr, E.d_grad, (const uint32_t *)E.d_rowmap[i], (long long)n, CFDP_DIM2, E.s_comp);
  CUDA_CHECK(cudaMemcpyAsync(&sd->grad[0][0][0], E.d_stage, n * CFDP_DIM2 * sizeof(double), cudaMemcpyDeviceToHost, E.s_comp));
  CUDA_CHECK(cudaStreamSynchronize(E.s_comp));
}

/* sd->grad (all rows, ghosts included) -> device: what compute_psd_flux reads when the caller keeps its state on the host */
extern "C" void cfdp_grad_to_device(solver_data *sd)
{
  Engine &E = g_eng;
  cfdp_commit();
  Domain *d = engine_find_domain(sd);
  ASSERT(d != NULL);
  const int i = hosted_index(d);
  const size_t n = (size_t)sd->nallpoints;
  ASSERT(sd->grad != NULL); /* CFDP_LEAN_HOST: there is no host mirror of grad */
  CUDA_CHECK(cudaMemcpyAsync(E.d_stage, &sd->grad[0][0][0], n * CFDP_DIM2 * sizeof(double), cudaMemcpyHostToDevice, E.s_comp));
  launch_rows_copy(E.d_grad, (const uint32_t *)E.d_rowmap[i], E.d_stage, nullptr, (long long)n, CFDP_DIM2, E.s_comp);
}

/* device pseudo flux -> own rows of sd->psd_flux (ghost rows are not defined, flux_kernels.cuh) */
extern "C" void cfdp_flux_to_host(solver_data *sd)
{
  Engine &E = g_eng;
  ASSERT(E.committed);
  flux_prepare();
  Domain *d = engine_find_domain(sd);
  ASSERT(d != NULL);
  const int i = hosted_index(d);
  const size_t n = (size_t)sd->nownpoints;
  ASSERT(sd->psd_flux != NULL); /* CFDP_LEAN_HOST: there is no host mirror of psd_flux */
  launch_rows_copy(E.d_stage, nullptr, E.d_flux, (const uint32_t *)E.d_rowmap[i], (long long)n, NFLUX, E.s_comp);
  CUDA_CHECK(cudaMemcpyAsync(&sd->psd_flux[0][0], E.d_stage, n * NFLUX * sizeof(double), cudaMemcpyDeviceToHost, E.s_comp));
  CUDA_CHECK(cudaStreamSynchronize(E.s_comp));
}

/* grad rows of the host container that the exchange overwrites must be on the device before
 * a drop-in call when the caller modified them: not needed, ghosts are fully rewritten */

/* ------------------------------------------------------------------------------------------
 * one iteration = gradient of all hosted domains + halo exchange of grad
 * ---------------------------------------------------------------------------------------- */
static bool have_exchange(void) { return g_eng.n_local > 0 || !g_eng.peers.empty(); }

/* ------------------------------------------------------------------------------------------
 * One-sided backend: put + notify over CUDA IPC (the GASPI write_notify / MPI_Put model,
 * exchange_data_gaspi.c:105-151, exchange_data_mpidma.c:93-127).  Every rank owns a receive window of two
 * stages (double buffered by stage parity like the reference's segments, exchange_data_gaspi.c:181,230) and one
 * arrival counter per peer.  A sender copies its packed rows straight into the peer's window at the offset the
 * peer told it at setup (remote_recv_offset, comm_data.c:355-396) -- a device-to-device copy over NVLink that
 * needs no SM -- and then bumps its counter in the peer's memory; the receiver's stream sleeps on its own
 * counter (stream memory operation), then unpacks.  No kernel ever spins.
 * ---------------------------------------------------------------------------------------- */
__global__ void notify_kernel(unsigned long long *flag, unsigned long long value)
{
  __threadfence_system();
  *reinterpret_cast<volatile unsigned long long *>(flag) = value;
  __threadfence_system();
}

/* exchange one int with EVERY other process and AND the results: all ranks must agree on the transport, or one side
 * would wait on a flag while the other sits in an NCCL group */
static bool all_procs_agree(bool mine)
{
  Engine &E = g_eng;
  if (E.nprocs == 1) return mine;
  const int n = E.nprocs - 1;
  std::vector<int> sv((size_t)n, mine ? 1 : 0), rv((size_t)n, 0);
  std::vector<int> peer; std::vector<const int *> sp; std::vector<int> sc; std::vector<int *> rp; std::vector<int> rc;
  for (int q = 0, k = 0; q < E.nprocs; q++) if (q != E.proc_rank) { peer.push_back(q); sp.push_back(&sv[(size_t)k]); sc.push_back(1); rp.push_back(nullptr); rc.push_back(0); k++; }
  for (int q = 0, k = 0; q < E.nprocs; q++) if (q != E.proc_rank) { peer.push_back(q); sp.push_back(nullptr); sc.push_back(0); rp.push_back(&rv[(size_t)k]); rc.push_back(1); k++; }
  engine_exchange_ints(peer, sp, sc, rp, rc);
  bool all = mine;
  for (int v : rv) all = all && v != 0;
  return all;
}

static void ipc_close_peers(void)
{
  for (PeerPlan &p : g_eng.peers) {
    if (p.opened) {
      if (p.peer_recvbuf) cudaIpcCloseMemHandle(p.peer_recvbuf);
      if (p.peer_flags) cudaIpcCloseMemHandle(p.peer_flags);
      if (p.peer_grad) cudaIpcCloseMemHandle(p.peer_grad);
    }
    p.peer_recvbuf = nullptr; p.peer_flags = nullptr; p.peer_grad = nullptr; p.opened = false;
  }
}

static void ipc_setup(void)
{
  Engine &E = g_eng;
  if (E.ipc_ready || E.peers.empty()) return;
  const size_t np = E.peers.size();
  CUDA_CHECK(cudaMalloc(&E.d_recvwin, (size_t)std::max<long long>(E.n_recv, 1) * 2 * CFDP_DIM2 * sizeof(double)));
  CUDA_CHECK(cudaMalloc(&E.d_arrived, 256 * sizeof(unsigned long long)));
  CUDA_CHECK(cudaMemset(E.d_arrived, 0, 256 * sizeof(unsigned long long)));
  /* 1. handles of my receive window, counters and grad array + where each peer's rows go + my unpack rows for that
   * peer (they become the destinations of the peer's direct stores) */
  cudaIpcMemHandle_t hwin, hflag, hgrad;
  memset(&hwin, 0, sizeof hwin); memset(&hflag, 0, sizeof hflag); memset(&hgrad, 0, sizeof hgrad);
  if (!E.loopback) {
    CUDA_CHECK(cudaIpcGetMemHandle(&hwin, E.d_recvwin));
    CUDA_CHECK(cudaIpcGetMemHandle(&hflag, E.d_arrived));
    CUDA_CHECK(cudaIpcGetMemHandle(&hgrad, E.d_grad));
  }
  const int HW = (int)(sizeof(cudaIpcMemHandle_t) / sizeof(int)); /* 16 */
  const int HDR = 3 * HW + 5;
  std::vector<std::vector<int>> sb(np), rb(np);
  std::vector<int> peer; std::vector<const int *> sp; std::vector<int> sc; std::vector<int *> rp; std::vector<int> rc;
  for (size_t i = 0; i < np; i++) {
    const PeerPlan &p = E.peers[i];
    sb[i].assign((size_t)HDR + (size_t)p.recv_rows, 0);
    memcpy(sb[i].data(), &hwin, sizeof hwin); memcpy(sb[i].data() + HW, &hflag, sizeof hflag); memcpy(sb[i].data() + 2 * HW, &hgrad, sizeof hgrad);
    sb[i][3 * HW + 0] = (int)(p.recv_off & 0x7FFFFFFF); sb[i][3 * HW + 1] = (int)(p.recv_off >> 31);
    sb[i][3 * HW + 2] = (int)i;                                   /* the counters this peer bumps */
    sb[i][3 * HW + 3] = (int)(E.n_recv & 0x7FFFFFFF); sb[i][3 * HW + 4] = (int)(E.n_recv >> 31);
    for (long long j = 0; j < p.recv_rows; j++) sb[i][(size_t)HDR + (size_t)j] = (int)E.h_recv_rows[(size_t)(p.recv_off + j)];
    rb[i].assign((size_t)HDR + (size_t)p.send_rows, 0);           /* the peer receives what I send */
    peer.push_back(p.proc); sp.push_back(sb[i].data()); sc.push_back((int)sb[i].size()); rp.push_back(nullptr); rc.push_back(0);
  }
  for (size_t i = 0; i < np; i++) { peer.push_back(E.peers[i].proc); sp.push_back(nullptr); sc.push_back(0); rp.push_back(rb[i].data()); rc.push_back((int)rb[i].size()); }
  engine_exchange_ints(peer, sp, sc, rp, rc);
  /* 2. map the peers' buffers */
  bool ok = get_wait_value64() != nullptr;
  for (size_t i = 0; i < np && ok; i++) {
    PeerPlan &p = E.peers[i];
    if (E.loopback) { p.peer_recvbuf = E.d_recvwin; p.peer_flags = E.d_arrived; p.peer_grad = E.d_grad; }
    else {
      cudaIpcMemHandle_t h1, h2, h3;
      memcpy(&h1, rb[i].data(), sizeof h1); memcpy(&h2, rb[i].data() + HW, sizeof h2); memcpy(&h3, rb[i].data() + 2 * HW, sizeof h3);
      void *w = nullptr, *f = nullptr, *g = nullptr;
      p.opened = true;
      if (cudaIpcOpenMemHandle(&w, h1, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
          cudaIpcOpenMemHandle(&f, h2, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
          cudaIpcOpenMemHandle(&g, h3, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        fprintf(stderr, "cfdp: rank %d cannot map the buffers of rank %d over CUDA IPC (%s)\n", E.proc_rank, p.proc, cudaGetErrorString(cudaGetLastError()));
        ok = false;
      }
      p.peer_recvbuf = (double *)w; p.peer_flags = (unsigned long long *)f; p.peer_grad = (double *)g;
    }
    p.remote_recv_off = (long long)rb[i][3 * HW + 0] | ((long long)rb[i][3 * HW + 1] << 31);
    p.remote_slot = rb[i][3 * HW + 2];
    p.remote_recv_total = (long long)rb[i][3 * HW + 3] | ((long long)rb[i][3 * HW + 4] << 31);
    p.peer_rows.resize((size_t)p.send_rows);
    for (long long j = 0; j < p.send_rows; j++) p.peer_rows[(size_t)j] = (uint32_t)rb[i][(size_t)HDR + (size_t)j];
  }
  /* 3. every rank uses the one-sided transports, or none does (a rank that fell back to NCCL alone would
   * deadlock against peers waiting on its flags) */
  if (!all_procs_agree(ok)) {
    if (E.proc_rank == 0) fprintf(stderr, "cfdp: CUDA IPC peer mapping is not available on every rank: the one-sided variants (gaspi_*, mpifence_*, mpipscw_*) use the NCCL transport\n");
    ipc_close_peers();
    return;
  }
  E.ipc_ready = true;

  /* 4. direct halo stores: the export lists once more, remote rows addressed in the peers' grad arrays
   * (destination array 2 + peer), and per boundary tile the (peer, rows) pairs its stores complete */
  {
    std::vector<uint32_t> src2(E.h_exp_src), dst2(E.h_exp_dst);
    std::vector<std::vector<uint32_t>> tile_peer_rows((size_t)E.nbtiles, std::vector<uint32_t>(np, 0));
    for (long long t = 0; t < E.nbtiles; t++)
      for (uint32_t e = E.h_exp_off[(size_t)t]; e < E.h_exp_off[(size_t)t + 1]; e++) {
        if ((src2[e] >> 16) != 1u) continue;                       /* a slot of the packed send buffer ... */
        const long long slot = (long long)dst2[e];
        size_t pi = 0;
        while (pi < np && !(slot >= E.peers[pi].send_off && slot < E.peers[pi].send_off + E.peers[pi].send_rows)) pi++;
        ASSERT(pi < np);
        dst2[e] = E.peers[pi].peer_rows[(size_t)(slot - E.peers[pi].send_off)];   /* ... becomes a ghost row of the peer */
        src2[e] = (src2[e] & 0xFFFFu) | ((uint32_t)(2 + pi) << 16);
        tile_peer_rows[(size_t)t][pi]++;
      }
    std::vector<uint32_t> sig_off((size_t)E.nbtiles + 1, 0), sig_ent;
    for (long long t = 0; t < E.nbtiles; t++) {
      for (size_t pi = 0; pi < np; pi++)
        if (tile_peer_rows[(size_t)t][pi]) sig_ent.push_back((uint32_t)pi | (tile_peer_rows[(size_t)t][pi] << 4));
      sig_off[(size_t)t + 1] = (uint32_t)sig_ent.size();
    }
    E.d_exp_src_direct = upload(src2); E.d_exp_dst_direct = upload(dst2);
    E.d_sig_off = upload(sig_off); E.d_sig_ent = upload(sig_ent);
    for (size_t pi = 0; pi < np; pi++) {
      E.pipe.exp_base[2 + pi] = E.peers[pi].peer_grad;
      E.pipe.sig_flag[pi] = E.peers[pi].peer_flags + FLAG_ROWS + E.peers[pi].remote_slot;
    }
    E.direct_ready = env_int("CFDP_DIRECT", 1) != 0;
    E.direct_epoch = 0;
  }
}

static CUresult stream_wait_geq(cudaStream_t st, const void *addr, unsigned long long value);

static void enqueue_exchange_onesided(cudaStream_t st, bool packed)
{
  Engine &E = g_eng;
  if (!packed) launch_rows_copy(E.d_grad, E.d_loc_dst, E.d_grad, E.d_loc_src, E.n_local, CFDP_DIM2, st);
  E.last_transport = 1;
  if (E.peers.empty()) return;
  E.last_transport = 3;
  const unsigned long long stage = E.ipc_stage++;
  const int half = (int)(stage & 1);                                                               /* exchange_data_gaspi.c:181 */
  if (!packed) launch_rows_copy(E.d_sendbuf, nullptr, E.d_grad, E.d_send_rows, E.n_send, CFDP_DIM2, st);      /* threads.c:791-813 */
  for (const PeerPlan &p : E.peers) {
    if (!p.send_rows) continue;
    double *dst = p.peer_recvbuf + ((size_t)half * (size_t)p.remote_recv_total + (size_t)p.remote_recv_off) * CFDP_DIM2;
    CUDA_CHECK(cudaMemcpyAsync(dst, E.d_sendbuf + p.send_off * CFDP_DIM2, (size_t)p.send_rows * CFDP_DIM2 * sizeof(double),
                               cudaMemcpyDeviceToDevice, st));                                      /* gaspi_write ... */
    notify_kernel<<<1, 1, 0, st>>>(p.peer_flags + FLAG_STAGE + p.remote_slot, stage + 1);           /* ... _notify */
    CUDA_CHECK(cudaGetLastError());
    E.launches++;
  }
  for (size_t i = 0; i < E.peers.size(); i++) {                                                     /* gaspi_notify_waitsome, exchange_data_gaspi.c:252-262 */
    if (!E.peers[i].recv_rows) continue;
    CUresult r = stream_wait_geq(st, E.d_arrived + FLAG_STAGE + i, stage + 1);
    ASSERT(r == CUDA_SUCCESS);
  }
  launch_rows_copy(E.d_grad, E.d_recv_rows, E.d_recvwin + (size_t)half * (size_t)E.n_recv * CFDP_DIM2, nullptr, E.n_recv, CFDP_DIM2, st);
}

static void enqueue_exchange(cudaStream_t st, bool packed)
{
  Engine &E = g_eng;
  /* halo rows whose owner lives on this GPU: one gather/scatter, no staging */
  if (!packed) launch_rows_copy(E.d_grad, E.d_loc_dst, E.d_grad, E.d_loc_src, E.n_local, CFDP_DIM2, st);
  E.last_transport = 1;
  if (E.peers.empty()) return;
  E.last_transport = 2;
  if (!packed) launch_rows_copy(E.d_sendbuf, nullptr, E.d_grad, E.d_send_rows, E.n_send, CFDP_DIM2, st);      /* threads.c:791-813 */
  if (E.loopback) { /* test mode: this process is its own peer, the message is one device copy */
    ASSERT(E.n_send == E.n_recv);
    CUDA_CHECK(cudaMemcpyAsync(E.d_recvbuf, E.d_sendbuf, (size_t)E.n_send * CFDP_DIM2 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    launch_rows_copy(E.d_grad, E.d_recv_rows, E.d_recvbuf, nullptr, E.n_recv, CFDP_DIM2, st);
    return;
  }
  NCCL_CHECK(g_nccl.GroupStart());
  for (const PeerPlan &p : E.peers) {                                                               /* exchange_data_mpi.c:96-166 */
    if (p.recv_rows) NCCL_CHECK(g_nccl.Recv(E.d_recvbuf + p.recv_off * CFDP_DIM2, (size_t)p.recv_rows * CFDP_DIM2, NCCL_FLOAT64, p.proc, E.comm, st));
    if (p.send_rows) NCCL_CHECK(g_nccl.Send(E.d_sendbuf + p.send_off * CFDP_DIM2, (size_t)p.send_rows * CFDP_DIM2, NCCL_FLOAT64, p.proc, E.comm, st));
  }
  NCCL_CHECK(g_nccl.GroupEnd());
  launch_rows_copy(E.d_grad, E.d_recv_rows, E.d_recvbuf, nullptr, E.n_recv, CFDP_DIM2, st);      /* threads.c:816-839 */
}

static wait_value64_fn get_wait_value64(void)
{
  static wait_value64_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue64", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) fn = (wait_value64_fn)p;
    else (void)cudaGetLastError();
  }
  return fn;
}

static CUresult stream_wait_geq(cudaStream_t st, const void *addr, unsigned long long value)
{
  wait_value64_fn fn = get_wait_value64();
  if (!fn) return CUDA_ERROR_NOT_SUPPORTED;
  return fn((CUstream)st, (CUdeviceptr)addr, (cuuint64_t)value, CU_STREAM_WAIT_VALUE_GEQ);
}

/* packed: the gradient kernel already wrote the send buffer and the ghost rows of same-GPU partners (export lists) */
static void enqueue_exchange_for(int variant, cudaStream_t st, bool packed)
{
  Engine &E = g_eng;
  const bool onesided = (variant == CFDP_GASPI_BULK_SYNC || variant == CFDP_GASPI_ASYNC) && E.ipc_ready && get_wait_value64();
  if (onesided) enqueue_exchange_onesided(st, packed); else enqueue_exchange(st, packed);
}
static bool kernel_packs(void) { return g_eng.fused_pack && g_eng.kernel_version == 2; }

/* ------------------------------------------------------------------------------------------
 * Direct halo stores (the *_async one-sided variants): the boundary tiles of the gradient kernel store the rows a
 * peer GPU needs straight into the ghost rows of the peer's grad array (CUDA IPC mapping, NVLink) and add the number
 * of rows to the peer's arrival counter (red.release.sys) -- per partner, the moment a tile's send points are final
 * (threads.c:268-306, exchange_data_gaspi.c:105-151).  No pack, no transfer, no unpack kernel.  A rank's iteration
 * is complete when every peer's counter has reached epoch * rows (stream wait, no kernel spins).  Before a rank
 * overwrites the ghost rows of a peer in the next epoch it needs that peer's credit: "every consumer of the previous
 * rows that was enqueued before my next iteration has finished" (the analogue of the reference's double-buffered
 * segments + queue back-pressure, exchange_data_gaspi.c:181,230, queue.c:20-33).
 * ---------------------------------------------------------------------------------------- */
struct CreditArgs { unsigned long long *flag[MAX_PEERS]; int n; };
__global__ void credit_kernel(CreditArgs a, unsigned long long value)
{
  const int i = threadIdx.x;
  if (i < a.n) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long *>(a.flag[i]) = value;
    __threadfence_system();
  }
}

static bool use_direct(int variant) { return variant == CFDP_GASPI_ASYNC && g_eng.direct_ready && kernel_packs() && !g_eng.peers.empty(); }

static void run_iteration_direct(void)
{
  Engine &E = g_eng;
  const unsigned long long e = ++E.direct_epoch;
  if (e > 1) {
    CreditArgs ca; ca.n = 0;
    for (const PeerPlan &p : E.peers) if (p.recv_rows) ca.flag[ca.n++] = p.peer_flags + FLAG_CREDIT + p.remote_slot;
    if (ca.n) { credit_kernel<<<1, 32, 0, E.s_comp>>>(ca, e - 1); CUDA_CHECK(cudaGetLastError()); E.launches++; }
    for (size_t i = 0; i < E.peers.size(); i++)
      if (E.peers[i].send_rows) { CUresult r = stream_wait_geq(E.s_comp, E.d_arrived + FLAG_CREDIT + i, e - 1); ASSERT(r == CUDA_SUCCESS); }
  }
  launch_gradient(0, E.ntiles, E.s_comp, 0, true, true);
  if (E.timeline_ek) CUDA_CHECK(cudaEventRecord(E.timeline_ek, E.s_comp));
  for (size_t i = 0; i < E.peers.size(); i++)
    if (E.peers[i].recv_rows) { CUresult r = stream_wait_geq(E.s_comp, E.d_arrived + FLAG_ROWS + i, e * (unsigned long long)E.peers[i].recv_rows); ASSERT(r == CUDA_SUCCESS); }
  if (E.timeline_ex) CUDA_CHECK(cudaEventRecord(E.timeline_ex, E.s_comp));
  E.last_transport = 4;
  for (Domain *d : E.doms)
    if (d->cd->ndomains > 1) { d->cd->send_stage++; d->cd->recv_stage++; d->cd->comm_stage++; }
}

static void run_iteration(int variant)
{
  Engine &E = g_eng;
  const bool overlap = (variant == CFDP_MPI_ASYNC || variant == CFDP_GASPI_ASYNC);
  if (variant != CFDP_COMM_FREE && use_direct(variant)) { run_iteration_direct(); return; }
  if (variant == CFDP_COMM_FREE || !have_exchange()) {
    launch_gradient(0, E.ntiles, E.s_comp);                       /* gradients.c:150-165 */
    E.last_transport = 0;
  } else if (!overlap) {
    launch_gradient(0, E.ntiles, E.s_comp, 0, true);              /* bulk synchronous: compute (packing on the way, threads.c:187-249), then exchange (exchange_data_mpi.c:199-284) */
    if (E.timeline_ek) CUDA_CHECK(cudaEventRecord(E.timeline_ek, E.s_comp));
    enqueue_exchange_for(variant, E.s_comp, kernel_packs());
    if (E.timeline_ex) CUDA_CHECK(cudaEventRecord(E.timeline_ex, E.s_comp));
  } else {
    /* early send (threads.c:253-346): tiles holding send points first, their rows are packed and
     * shipped on the comm stream while the interior tiles compute */
    wait_value64_fn wait64 = (E.fused_signal && E.kernel_version == 2) ? get_wait_value64() : nullptr;
    if (wait64) {
      /* ONE launch over all tiles; every retired boundary tile bumps a device counter and the comm stream
       * sleeps on the counter (stream memory operation) -- no split launch, no tail between the two parts */
      E.progress_target += (unsigned long long)E.nbtiles;
      launch_gradient(0, E.ntiles, E.s_comp, (int)E.nbtiles, true);
      CUresult r = wait64((CUstream)E.s_comm, (CUdeviceptr)E.d_progress, (cuuint64_t)E.progress_target, CU_STREAM_WAIT_VALUE_GEQ);
      ASSERT(r == CUDA_SUCCESS);
      if (E.timeline_ek) CUDA_CHECK(cudaEventRecord(E.timeline_ek, E.s_comp));
      enqueue_exchange_for(variant, E.s_comm, kernel_packs());
      if (E.timeline_ex) CUDA_CHECK(cudaEventRecord(E.timeline_ex, E.s_comm));
      CUDA_CHECK(cudaEventRecord(E.ev_x, E.s_comm));
      CUDA_CHECK(cudaStreamWaitEvent(E.s_comp, E.ev_x, 0));
      for (Domain *d : E.doms)
        if (d->cd->ndomains > 1) { d->cd->send_stage++; d->cd->recv_stage++; d->cd->comm_stage++; }
      return;
    }
    launch_gradient(0, E.nbtiles, E.s_comp, 0, true);
    CUDA_CHECK(cudaEventRecord(E.ev_b, E.s_comp));
    CUDA_CHECK(cudaStreamWaitEvent(E.s_comm, E.ev_b, 0));
    enqueue_exchange_for(variant, E.s_comm, kernel_packs());
    if (E.timeline_ex) CUDA_CHECK(cudaEventRecord(E.timeline_ex, E.s_comm));
    CUDA_CHECK(cudaEventRecord(E.ev_x, E.s_comm));
    launch_gradient(E.nbtiles, E.ntiles - E.nbtiles, E.s_comp);
    if (E.timeline_ek) CUDA_CHECK(cudaEventRecord(E.timeline_ek, E.s_comp));
    CUDA_CHECK(cudaStreamWaitEvent(E.s_comp, E.ev_x, 0));        /* exchange_dbl_mpi_async waits for all partners before returning */
  }
  for (Domain *d : E.doms)
    if (d->cd->ndomains > 1 && variant != CFDP_COMM_FREE) { d->cd->send_stage++; d->cd->recv_stage++; d->cd->comm_stage++; }
}

/* debug (CFDP_TIMELINE=1): when, relative to the start of one iteration, the gradient kernel(s) and the
 * exchange finish */
static void timeline_probe(int variant)
{
  Engine &E = g_eng;
  cudaEvent_t e0, ek, ex;
  CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&ek)); CUDA_CHECK(cudaEventCreate(&ex));
  CUDA_CHECK(cudaDeviceSynchronize());
  CUDA_CHECK(cudaEventRecord(e0, E.s_comp));
  CUDA_CHECK(cudaStreamWaitEvent(E.s_comm, e0, 0));
  E.timeline_ek = ek; E.timeline_ex = ex;
  run_iteration(variant);
  E.timeline_ek = nullptr; E.timeline_ex = nullptr;
  CUDA_CHECK(cudaDeviceSynchronize());
  float tk = 0, tx = 0;
  CUDA_CHECK(cudaEventElapsedTime(&tk, e0, ek));
  if (cudaEventElapsedTime(&tx, e0, ex) != cudaSuccess) { (void)cudaGetLastError(); tx = -1; }
  fprintf(stderr, "cfdp timeline (variant %d, rank %d): gradient kernel(s) done at %.3f ms, exchange done at %.3f ms\n", variant, E.proc_rank, tk, tx);
  cudaEventDestroy(e0); cudaEventDestroy(ek); cudaEventDestroy(ex);
}

extern "C" double cfdp_iterate(int variant, int niter, int final_last)
{
  (void)final_last; /* `final` only suppresses re-posting receives in the reference (exchange_data_mpi.c:527-531) */
  Engine &E = g_eng;
  cfdp_commit();
  ASSERT(variant >= CFDP_COMM_FREE && variant <= CFDP_GASPI_ASYNC);
  if (env_int("CFDP_TIMELINE", 0) && variant != CFDP_COMM_FREE) timeline_probe(variant);
  CUDA_CHECK(cudaEventRecord(E.ev_t0, E.s_comp));
  for (int i = 0; i < niter; i++) {
    if (E.var_refresh) launch_halo_pack(0, E.ntiles, E.s_comp);
    run_iteration(variant);
    if (E.with_flux) launch_flux(0, E.ntiles, E.s_comp); /* solver.c:45-55: gradient (+ exchange), then the pseudo flux; s_comp has joined the exchange */
  }
  CUDA_CHECK(cudaEventRecord(E.ev_t1, E.s_comp));
  CUDA_CHECK(cudaEventSynchronize(E.ev_t1));
  float ms = 0;
  CUDA_CHECK(cudaEventElapsedTime(&ms, E.ev_t0, E.ev_t1));
  if (variant == CFDP_COMM_FREE && niter > 0) E.last_kernel_ms = ms / niter;
  return (double)ms;
}

extern "C" void cfdp_set_flux(int on) { g_eng.with_flux = on ? 1 : 0; }
extern "C" void cfdp_set_var_refresh(int on) { g_eng.var_refresh = on ? 1 : 0; }

/* rebuild the packed halo rows of every tile from the device var rows (var was changed on the device); device time in ms */
extern "C" double cfdp_refresh_var(int niter)
{
  Engine &E = g_eng;
  cfdp_commit();
  CUDA_CHECK(cudaEventRecord(E.ev_t0, E.s_comp));
  for (int i = 0; i < niter; i++) launch_halo_pack(0, E.ntiles, E.s_comp);
  CUDA_CHECK(cudaEventRecord(E.ev_t1, E.s_comp));
  CUDA_CHECK(cudaEventSynchronize(E.ev_t1));
  float ms = 0;
  CUDA_CHECK(cudaEventElapsedTime(&ms, E.ev_t0, E.ev_t1));
  return (double)ms;
}

/* niter pseudo-flux passes over all hosted domains on the device grad as it stands; returns the device time in ms */
extern "C" double cfdp_flux_iterate(int niter)
{
  Engine &E = g_eng;
  cfdp_commit();
  CUDA_CHECK(cudaEventRecord(E.ev_t0, E.s_comp));
  for (int i = 0; i < niter; i++) launch_flux(0, E.ntiles, E.s_comp);
  CUDA_CHECK(cudaEventRecord(E.ev_t1, E.s_comp));
  CUDA_CHECK(cudaEventSynchronize(E.ev_t1));
  float ms = 0;
  CUDA_CHECK(cudaEventElapsedTime(&ms, E.ev_t0, E.ev_t1));
  if (niter > 0) E.last_flux_ms = ms / niter;
  return (double)ms;
}

/*
 * One drop-in step with HOST buffers: sd->var in, sd->grad out (the reference's calling convention).  PCIe is the
 * bottleneck (56 B up, 168 B down per point), so the hosted domains are pipelined over three streams: while domain
 * d+1 uploads, domain d computes and the own rows of domain d-1 go back; both PCIe directions are busy at once.
 * Ghost rows follow after the exchange: they are contiguous at the end of both the host array and the domain's
 * device rows, so they need no gather.
 */
struct E2EResources {
  bool ready = false;
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  double *var_stage[2] = {nullptr, nullptr}, *grad_stage[2] = {nullptr, nullptr};
  cudaEvent_t ev_up[2], ev_var_free[2], ev_gath[2], ev_grad_free[2], ev_begin, ev_x, ev_end;
};
static E2EResources g_e2e;

static void e2e_setup(void)
{
  Engine &E = g_eng;
  if (g_e2e.ready) return;
  size_t max_var = 0, max_grad = 0;
  for (Domain *d : E.doms) {
    max_var = std::max(max_var, (size_t)d->sch.nall * NGRAD * sizeof(double));
    max_grad = std::max(max_grad, (size_t)d->sch.nown * CFDP_DIM2 * sizeof(double));
  }
  CUDA_CHECK(cudaStreamCreateWithFlags(&g_e2e.s_h2d, cudaStreamNonBlocking));
  CUDA_CHECK(cudaStreamCreateWithFlags(&g_e2e.s_d2h, cudaStreamNonBlocking));
  for (int b = 0; b < 2; b++) {
    CUDA_CHECK(cudaMalloc(&g_e2e.var_stage[b], max_var));
    CUDA_CHECK(cudaMalloc(&g_e2e.grad_stage[b], max_grad));
    CUDA_CHECK(cudaEventCreateWithFlags(&g_e2e.ev_up[b], cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&g_e2e.ev_var_free[b], cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&g_e2e.ev_gath[b], cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&g_e2e.ev_grad_free[b], cudaEventDisableTiming));
  }
  CUDA_CHECK(cudaEventCreateWithFlags(&g_e2e.ev_begin, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreateWithFlags(&g_e2e.ev_x, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreateWithFlags(&g_e2e.ev_end, cudaEventDisableTiming));
  g_e2e.ready = true;
}

static void e2e_release(void)
{
  if (!g_e2e.ready) return;
  cudaStreamDestroy(g_e2e.s_h2d); cudaStreamDestroy(g_e2e.s_d2h);
  for (int b = 0; b < 2; b++) {
    cudaFree(g_e2e.var_stage[b]); cudaFree(g_e2e.grad_stage[b]);
    cudaEventDestroy(g_e2e.ev_up[b]); cudaEventDestroy(g_e2e.ev_var_free[b]); cudaEventDestroy(g_e2e.ev_gath[b]); cudaEventDestroy(g_e2e.ev_grad_free[b]);
  }
  cudaEventDestroy(g_e2e.ev_begin); cudaEventDestroy(g_e2e.ev_x); cudaEventDestroy(g_e2e.ev_end);
  g_e2e = E2EResources();
}

/* enqueue one pipelined host-buffer step; the caller synchronises on s_comp */
static void enqueue_step_e2e(int variant)
{
  Engine &E = g_eng;
  for (Domain *d : E.doms) ASSERT(d->sd->grad != NULL); /* host buffers in and out: not with CFDP_LEAN_HOST */
  e2e_setup();
  E2EResources &R = g_e2e;
  const bool exchange = variant != CFDP_COMM_FREE && have_exchange();
  CUDA_CHECK(cudaEventRecord(R.ev_begin, E.s_comp));
  CUDA_CHECK(cudaStreamWaitEvent(R.s_h2d, R.ev_begin, 0));
  CUDA_CHECK(cudaStreamWaitEvent(R.s_d2h, R.ev_begin, 0));
  const int nh = (int)E.doms.size();
  for (int i = 0; i < nh; i++) {
    Domain *d = E.doms[(size_t)i];
    solver_data *sd = d->sd;
    const int b = i & 1;
    const size_t nall = (size_t)sd->nallpoints, nown = (size_t)sd->nownpoints;
    /* up: var of this domain (own + ghost rows, host order) */
    if (i >= 2) CUDA_CHECK(cudaStreamWaitEvent(R.s_h2d, R.ev_var_free[b], 0));
    CUDA_CHECK(cudaMemcpyAsync(R.var_stage[b], &sd->var[0][0], nall * NGRAD * sizeof(double), cudaMemcpyHostToDevice, R.s_h2d));
    CUDA_CHECK(cudaEventRecord(R.ev_up[b], R.s_h2d));
    /* compute: permute into device rows (halved), this domain's boundary and interior tiles */
    CUDA_CHECK(cudaStreamWaitEvent(E.s_comp, R.ev_up[b], 0));
    launch_rows_copy(E.d_var, (const uint32_t *)E.d_rowmap[(size_t)i], R.var_stage[b], nullptr, (long long)nall, NGRAD, E.s_comp, 0.5);
    launch_halo_pack_domain(d, E.s_comp);
    CUDA_CHECK(cudaEventRecord(R.ev_var_free[b], E.s_comp));
    launch_gradient(d->tile0_b, d->sch.nboundary, E.s_comp, 0, exchange);
    launch_gradient(d->tile0_i, d->sch.ntiles - d->sch.nboundary, E.s_comp);
    /* down: own rows back in host order */
    if (i >= 2) CUDA_CHECK(cudaStreamWaitEvent(E.s_comp, R.ev_grad_free[b], 0));
    launch_rows_copy(R.grad_stage[b], nullptr, E.d_grad, (const uint32_t *)E.d_rowmap[(size_t)i], (long long)nown, CFDP_DIM2, E.s_comp);
    CUDA_CHECK(cudaEventRecord(R.ev_gath[b], E.s_comp));
    CUDA_CHECK(cudaStreamWaitEvent(R.s_d2h, R.ev_gath[b], 0));
    CUDA_CHECK(cudaMemcpyAsync(&sd->grad[0][0][0], R.grad_stage[b], nown * CFDP_DIM2 * sizeof(double), cudaMemcpyDeviceToHost, R.s_d2h));
    CUDA_CHECK(cudaEventRecord(R.ev_grad_free[b], R.s_d2h));
  }
  if (exchange) {
    /* every domain's rows are final: halo exchange, then the ghost rows (contiguous on both sides) */
    enqueue_exchange_for(variant, E.s_comp, kernel_packs());
    CUDA_CHECK(cudaEventRecord(R.ev_x, E.s_comp));
    CUDA_CHECK(cudaStreamWaitEvent(R.s_d2h, R.ev_x, 0));
    for (int i = 0; i < nh; i++) {
      Domain *d = E.doms[(size_t)i];
      solver_data *sd = d->sd;
      const size_t nadd = (size_t)(sd->nallpoints - sd->nownpoints);
      if (!nadd) continue;
      CUDA_CHECK(cudaMemcpyAsync(&sd->grad[sd->nownpoints][0][0], E.d_grad + (size_t)(d->rowbase + d->sch.ghost_row0) * CFDP_DIM2,
                                 nadd * CFDP_DIM2 * sizeof(double), cudaMemcpyDeviceToHost, R.s_d2h));
    }
    for (Domain *d : E.doms)
      if (d->cd->ndomains > 1) { d->cd->send_stage++; d->cd->recv_stage++; d->cd->comm_stage++; }
  }
  CUDA_CHECK(cudaEventRecord(R.ev_end, R.s_d2h));
  CUDA_CHECK(cudaStreamWaitEvent(E.s_comp, R.ev_end, 0));
}

extern "C" double cfdp_step_e2e(int variant)
{
  Engine &E = g_eng;
  cfdp_commit();
  CUDA_CHECK(cudaEventRecord(E.ev_t0, E.s_comp));
  enqueue_step_e2e(variant);
  CUDA_CHECK(cudaEventRecord(E.ev_t1, E.s_comp));
  CUDA_CHECK(cudaEventSynchronize(E.ev_t1));
  float ms = 0;
  CUDA_CHECK(cudaEventElapsedTime(&ms, E.ev_t0, E.ev_t1));
  return (double)ms;
}

/* debug: per-phase SM cycles of thread 0, summed over tiles (CFDP_PHASE_PROF=1) */
extern "C" int cfdp_get_phase_profile(unsigned long long *out8, int reset)
{
  Engine &E = g_eng;
  if (!E.pipe.prof) return -1;
  CUDA_CHECK(cudaDeviceSynchronize());
  CUDA_CHECK(cudaMemcpy(out8, E.pipe.prof, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (reset) CUDA_CHECK(cudaMemset(E.pipe.prof, 0, 8 * sizeof(unsigned long long)));
  return 0;
}

extern "C" void cfdp_device_synchronize(void)
{
  if (!g_eng.have_device) return;
  CUDA_CHECK(cudaStreamSynchronize(g_eng.s_comp));
  CUDA_CHECK(cudaStreamSynchronize(g_eng.s_comm));
}

/* the reference-named per-domain entry points.  They may be called by every OpenMP thread of a
 * parallel region (solver.c:45-55): only thread 0 acts. */
static void gradients_entry(comm_data *cd, solver_data *sd, int variant, int final)
{
  ASSERT(cd != NULL);
  ASSERT(sd != NULL);
  if (omp_get_thread_num() != 0) return;
  Engine &E = g_eng;
  Domain *d = engine_find_domain(sd);
  ASSERT(d != NULL && d->cd == cd);
  cfdp_commit();
  if (E.doms.size() != 1) {
    /* several hosted domains advance together: the call for the first hosted domain drives all */
    if (d != E.doms[0]) return;
  }
  (void)final;
  const int v = cd->ndomains == 1 ? CFDP_COMM_FREE : variant;
  if (E.resident) { run_iteration(v); return; }
  /* host buffers in, host buffers out (the reference's convention); returns when sd->grad is complete */
  enqueue_step_e2e(v);
  CUDA_CHECK(cudaStreamSynchronize(E.s_comp));
}

extern "C" void compute_gradients_gg_comm_free(comm_data *cd, solver_data *sd, int final) { gradients_entry(cd, sd, CFDP_COMM_FREE, final); }
extern "C" void compute_gradients_gg_mpi_bulk_sync(comm_data *cd, solver_data *sd, int final) { gradients_entry(cd, sd, CFDP_MPI_BULK_SYNC, final); }
extern "C" void compute_gradients_gg_mpi_early_recv(comm_data *cd, solver_data *sd, int final) { gradients_entry(cd, sd, CFDP_MPI_EARLY_RECV, final); }
extern "C" void compute_gradients_gg_mpi_async(comm_data *cd, solver_data *sd, int final) { gradients_entry(cd, sd, CFDP_MPI_ASYNC, final); }
extern "C" void compute_gradients_gg_gaspi_bulk_sync(comm_data *cd, solver_data *sd, int final) { gradients_entry(cd, sd, CFDP_GASPI_BULK_SYNC, final); }
extern "C" void compute_gradients_gg_gaspi_async(comm_data *cd, solver_data *sd, int final) { gradients_entry(cd, sd, CFDP_GASPI_ASYNC, final); }
/* one-sided MPI variants (USE_MPI_1_SIDED, exchange_data_mpidma.c): same write-then-signal data
 * flow as the GASPI variants on this backend */
extern "C" void compute_gradients_gg_mpifence_bulk_sync(comm_data *cd, solver_data *sd, int final) { gradients_entry(cd, sd, CFDP_GASPI_BULK_SYNC, final); }
extern "C" void compute_gradients_gg_mpifence_async(comm_data *cd, solver_data *sd, int final) { gradients_entry(cd, sd, CFDP_GASPI_ASYNC, final); }
extern "C" void compute_gradients_gg_mpipscw_bulk_sync(comm_data *cd, solver_data *sd, int final) { gradients_entry(cd, sd, CFDP_GASPI_BULK_SYNC, final); }
extern "C" void compute_gradients_gg_mpipscw_async(comm_data *cd, solver_data *sd, int final) { gradients_entry(cd, sd, CFDP_GASPI_ASYNC, final); }

/* flux.h:12.  Called by every OpenMP thread like the gradient (solver.c:52): only thread 0 acts; with several hosted
 * domains the call for the first one drives all.  Reads sd->grad (host) unless cfdp_set_resident(1), writes the own
 * rows of sd->psd_flux. */
extern "C" void compute_psd_flux(solver_data *sd)
{
  ASSERT(sd != NULL);
  if (omp_get_thread_num() != 0) return;
  Engine &E = g_eng;
  Domain *d = engine_find_domain(sd);
  ASSERT(d != NULL);
  cfdp_commit();
  if (d != E.doms[0]) return;
  if (!E.resident) for (Domain *h : E.doms) cfdp_grad_to_device(h->sd);
  launch_flux(0, E.ntiles, E.s_comp);
  if (!E.resident) for (Domain *h : E.doms) cfdp_flux_to_host(h->sd);
}

/* exchange_data_mpi.c:134-166 pre-posts MPI_Irecv; NCCL receives are enqueued together with the
 * sends, so this only validates its arguments */
extern "C" void exchange_dbl_mpi_post_recv(comm_data *cd, int dim2)
{
  ASSERT(cd != NULL);
  ASSERT(dim2 == CFDP_DIM2);
}

/* ------------------------------------------------------------------------------------------
 * introspection
 * ---------------------------------------------------------------------------------------- */
extern "C" void cfdp_get_stats(cfdp_stats *st)
{
  Engine &E = g_eng;
  memset(st, 0, sizeof *st);
  st->nprocs = E.configured ? E.nprocs : 0; st->proc_rank = E.proc_rank; st->ndomains_hosted = E.per_proc;
  st->tile_points = E.sopt.tile_points;
  if (!E.planned) return;
  st->nfaces = E.nfaces; st->nown = E.nown; st->nall = E.nall; st->rows = E.rows;
  st->ntiles = E.ntiles; st->nboundary_tiles = E.nbtiles; st->tile_faces = E.tile_faces; st->halo_refs = E.halo_refs;
  st->blob_bytes = (long long)E.blob_bytes; st->send_rows_local = E.n_local; st->send_rows_remote = E.n_send;
  st->alg_bytes = E.alg_bytes; st->h2d_bytes = E.nall * NGRAD * 8; st->d2h_bytes = E.nall * CFDP_DIM2 * 8;
  for (Domain *d : E.doms) { st->lds_wavefronts_min += d->sch.lds_wavefronts_min; st->lds_wavefronts_est += d->sch.lds_wavefronts_est; }
  st->launches = E.launches; st->last_kernel_ms = E.last_kernel_ms; st->smem_bytes = E.smem_bytes;
  st->halo_pack_bytes = E.halo_rows * NGRAD * 8; st->transport = E.last_transport; st->ipc_ready = E.ipc_ready ? 1 : 0; st->direct_ready = E.direct_ready ? 1 : 0; st->loopback = E.loopback ? 1 : 0;
  st->device_bytes = (long long)E.blob_bytes + (long long)E.fblob_bytes + E.rows * (NGRAD + CFDP_DIM2 + 1) * 8 + E.halo_rows * NGRAD * 8 + (long long)E.stage_bytes + (E.d_flux ? E.rows * NFLUX * 8 : 0) + (E.n_send + E.n_recv * 3) * CFDP_DIM2 * 8;
  st->flux_alg_bytes = E.flux_alg_bytes; st->last_flux_ms = E.last_flux_ms; st->flux_smem_bytes = (int)E.flux_smem; st->flux_blob_bytes = (long long)E.fblob_bytes;
}

extern "C" int cfdp_get_schedule(const solver_data *sd, cfdp_schedule_view *v)
{
  Domain *d = engine_find_domain(sd);
  if (!d || !g_eng.planned) return -1;
  const DomainSchedule &s = d->sch;
  v->ntiles = s.ntiles; v->nboundary_tiles = s.nboundary; v->nrows = s.nrows;
  v->row_of_point = s.row_of_point.data(); v->tile_row0 = s.tile_row0.data(); v->tile_npts = s.tile_npts.data();
  v->tile_nfaces = s.tile_nfaces.data(); v->tile_nhalo = s.tile_nhalo.data(); v->tile_is_boundary = s.tile_is_boundary.data();
  return 0;
}

extern "C" int cfdp_get_tile(const solver_data *sd, int tile, int *face_ids, int *halo_points)
{
  Domain *d = engine_find_domain(sd);
  if (!d || !g_eng.planned || tile < 0 || tile >= d->sch.ntiles) return -1;
  const DomainSchedule &s = d->sch;
  if (s.tile_face_ids.empty() && s.tile_faces > 0) return -1; /* released (CFDP_LEAN_HOST) */
  if (face_ids) memcpy(face_ids, &s.tile_face_ids[(size_t)s.tile_face_off[tile]], (size_t)s.tile_nfaces[tile] * sizeof(int));
  if (halo_points) memcpy(halo_points, &s.tile_halo_pts[(size_t)s.tile_halo_off[tile]], (size_t)s.tile_nhalo[tile] * sizeof(int));
  return s.tile_nfaces[tile];
}

static int rows_to_points(const Domain *d, const std::vector<uint32_t> &rows, int *points)
{
  const int i = hosted_index(d);
  for (size_t j = 0; j < rows.size(); j++) points[j] = g_eng.point_of_row[(size_t)i][(size_t)(rows[j] - d->rowbase)];
  return (int)rows.size();
}
extern "C" int cfdp_get_pack_list(const comm_data *cd, int partner, int *points)
{
  Domain *d = engine_find_domain(cd);
  if (!d || !g_eng.planned) return -1;
  auto it = g_eng.send_rows_of.find({d->id, partner});
  if (it == g_eng.send_rows_of.end()) return 0;
  return rows_to_points(d, it->second, points);
}
extern "C" int cfdp_get_unpack_list(const comm_data *cd, int partner, int *points)
{
  Domain *d = engine_find_domain(cd);
  if (!d || !g_eng.planned) return -1;
  auto it = g_eng.recv_rows_of.find({d->id, partner});
  if (it == g_eng.recv_rows_of.end()) return 0;
  return rows_to_points(d, it->second, points);
}
extern "C" int cfdp_get_sendbuf(const comm_data *cd, int partner, double *rows_out)
{
  Engine &E = g_eng;
  Domain *d = engine_find_domain(cd);
  if (!d || !E.committed) return -1;
  auto it = E.send_rows_of.find({d->id, partner});
  if (it == E.send_rows_of.end() || it->second.empty()) return 0;
  const size_t n = it->second.size();
  uint32_t *d_rows = upload(it->second);
  double *d_buf = nullptr;
  CUDA_CHECK(cudaMalloc(&d_buf, n * CFDP_DIM2 * sizeof(double)));
  CUDA_CHECK(cudaStreamSynchronize(E.s_comm));
  launch_rows_copy(d_buf, nullptr, E.d_grad, d_rows, (long long)n, CFDP_DIM2, E.s_comp); /* the pack kernel itself */
  CUDA_CHECK(cudaMemcpyAsync(rows_out, d_buf, n * CFDP_DIM2 * sizeof(double), cudaMemcpyDeviceToHost, E.s_comp));
  CUDA_CHECK(cudaStreamSynchronize(E.s_comp));
  CUDA_CHECK(cudaFree(d_rows)); CUDA_CHECK(cudaFree(d_buf));
  return (int)n;
}

extern "C" int cfdp_get_peer_plan(int i, int *proc, long long *send_rows, long long *recv_rows)
{
  Engine &E = g_eng;
  if (!E.planned) return -1;
  if (i < 0 || i >= (int)E.peers.size()) return (int)E.peers.size();
  *proc = E.peers[(size_t)i].proc; *send_rows = E.peers[(size_t)i].send_rows; *recv_rows = E.peers[(size_t)i].recv_rows;
  return (int)E.peers.size();
}

extern "C" int cfdp_get_exchange_entry(int dir, long long j, int *domain, int *point)
{
  Engine &E = g_eng;
  if (!E.planned) return -1;
  const std::vector<uint32_t> &rows = dir ? E.h_recv_rows : E.h_send_rows;
  if (j < 0 || j >= (long long)rows.size()) return -1;
  const long long r = rows[(size_t)j];
  for (size_t i = 0; i < E.doms.size(); i++) {
    Domain *d = E.doms[i];
    if (r >= d->rowbase && r < d->rowbase + d->sch.nrows) { *domain = d->id; *point = E.point_of_row[i][(size_t)(r - d->rowbase)]; return 0; }
  }
  return -1;
}

extern "C" int cfdp_get_row_owner(long long row, int *domain, int *point)
{
  Engine &E = g_eng;
  if (!E.planned) return -1;
  for (size_t i = 0; i < E.doms.size(); i++) {
    Domain *d = E.doms[i];
    if (row >= d->rowbase && row < d->rowbase + d->sch.nrows) {
      *domain = d->id; *point = E.point_of_row[i][(size_t)(row - d->rowbase)];
      return *point >= 0 ? 0 : -1;
    }
  }
  return -1;
}

/* raw tile blob for tests (host copy: available between cfdp_plan and cfdp_commit).  which = 0: gradient blob,
 * 1: pseudo-flux blob.  desc8 = {row0, npts, nhalo, nfaces | zslot << 16, maxdeg, npad, blob_bytes, halo_off}; halo rows are device rows. */
extern "C" long long cfdp_get_tile_blob(const solver_data *sd, int tile, int which, unsigned *desc8, unsigned char *bytes, long long capacity)
{
  Engine &E = g_eng;
  Domain *d = engine_find_domain(sd);
  if (!d || !E.planned || E.committed) return -1;
  const DomainSchedule &s = d->sch;
  if (tile < 0 || tile >= s.ntiles) return -1;
  if (which == 1 && !E.sopt.flux_blob) return -1;
  const long long slot = tile < s.nboundary ? d->tile0_b + tile : d->tile0_i + (tile - s.nboundary);
  const TileDesc &t = which ? E.h_ftiles[(size_t)slot] : E.h_tiles[(size_t)slot];
  const std::vector<unsigned char> &b = which ? s.fblob : s.blob;
  const uint64_t off = which ? s.ftile_blob[(size_t)tile] : s.tile_blob[(size_t)tile];
  if (desc8) { desc8[0] = t.row0; desc8[1] = t.npts; desc8[2] = t.nhalo; desc8[3] = (unsigned)t.nfaces | ((unsigned)t.zslot << 16); desc8[4] = t.maxdeg; desc8[5] = t.npad; desc8[6] = t.blob_bytes; desc8[7] = t.halo_off; }
  if (bytes && capacity >= (long long)t.blob_bytes) memcpy(bytes, &b[(size_t)off], t.blob_bytes);
  return (long long)t.blob_bytes;
}

extern "C" int cfdp_get_tile_exports(int tile, int capacity, unsigned *src_row, unsigned *dst, int *kind)
{
  Engine &E = g_eng;
  if (!E.planned || tile < 0 || tile >= E.nbtiles) return -1;
  const uint32_t e0 = E.h_exp_off[(size_t)tile], e1 = E.h_exp_off[(size_t)tile + 1];
  for (uint32_t e = e0; e < e1 && (int)(e - e0) < capacity; e++) {
    if (src_row) src_row[e - e0] = E.h_tiles[(size_t)tile].row0 + (E.h_exp_src[e] & 0xFFFFu);
    if (dst) dst[e - e0] = E.h_exp_dst[e];
    if (kind) kind[e - e0] = (int)(E.h_exp_src[e] >> 16);
  }
  return (int)(e1 - e0);
}

extern "C" void cfdp_finalize(void)
{
  Engine &E = g_eng;
  if (E.have_device) {
    cudaDeviceSynchronize();
    if (E.comm) { g_nccl.CommDestroy(E.comm); E.comm = nullptr; }
    cudaFree(E.d_var); cudaFree(E.d_grad); cudaFree(E.d_pvol); cudaFree(E.d_hhalo); E.d_hhalo = nullptr; cudaFree(E.d_flux); E.d_flux = nullptr; cudaFree(E.d_fblob); cudaFree(E.d_ftiles); E.d_fblob = nullptr; E.d_ftiles = nullptr; cudaFree(E.d_blob); cudaFree(E.d_tiles); cudaFree(E.d_stage);
    cudaFree(E.d_loc_dst); cudaFree(E.d_loc_src); cudaFree(E.d_exp_off); cudaFree(E.d_exp_src); cudaFree(E.d_exp_dst); cudaFree(E.d_send_rows); cudaFree(E.d_recv_rows); cudaFree(E.d_sendbuf); cudaFree(E.d_recvbuf);
    for (int *p : E.d_rowmap) cudaFree(p);
  }
  ipc_close_peers();
  cudaFree(E.d_recvwin); cudaFree(E.d_arrived); E.d_recvwin = nullptr; E.d_arrived = nullptr; E.ipc_ready = false; E.ipc_stage = 0;
  cudaFree(E.d_exp_src_direct); cudaFree(E.d_exp_dst_direct); cudaFree(E.d_sig_off); cudaFree(E.d_sig_ent);
  E.d_exp_src_direct = E.d_exp_dst_direct = E.d_sig_off = E.d_sig_ent = nullptr; E.direct_ready = false; E.direct_epoch = 0; E.last_transport = 0;
  for (Domain *d : E.doms) {
    /* host containers this library allocated (read_solver_data / read_communication_data / cfdp_attach_mesh) */
    if (d->sd) {
      solver_data *sd = d->sd;
      engine_free_pinned(sd->var); engine_free_pinned(sd->grad);
      free(sd->fpoint); free(sd->fnormal); free(sd->pvolume); free(sd->psd_flux);
      sd->var = nullptr; sd->grad = nullptr; sd->fpoint = nullptr; sd->fnormal = nullptr; sd->pvolume = nullptr; sd->psd_flux = nullptr;
    }
    if (d->cd && d->cd->ndomains > 1) {
      comm_data *cd = d->cd;
      if (cd->sendindex) for (int k = 0; k < cd->nProc; k++) { free(cd->sendindex[k]); free(cd->recvindex[k]); }
      free(cd->sendindex); free(cd->recvindex); free(cd->commpartner); free(cd->sendcount); free(cd->recvcount);
      free(cd->addpoint_owner); free(cd->addpoint_id); free(cd->local_recv_offset); free(cd->local_send_offset);
      free(cd->remote_recv_offset); free(cd->notification); free((void *)cd->recv_flag); free((void *)cd->send_flag);
      /* the struct is the caller's: only the pointers this library allocated are reset */
      cd->sendindex = cd->recvindex = nullptr; cd->commpartner = cd->sendcount = cd->recvcount = nullptr;
      cd->addpoint_owner = cd->addpoint_id = nullptr; cd->local_recv_offset = cd->local_send_offset = cd->remote_recv_offset = nullptr;
      cd->notification = nullptr; cd->recv_flag = cd->send_flag = nullptr;
    }
    delete d;
  }
  E.doms.clear(); E.d_rowmap.clear(); E.point_of_row.clear(); E.peers.clear(); E.send_rows_of.clear(); E.recv_rows_of.clear();
  if (E.pipe.prof) { cudaFree(E.pipe.prof); E.pipe.prof = nullptr; }
  if (E.d_progress) { cudaFree(E.d_progress); E.d_progress = nullptr; }
  e2e_release();
  if (E.have_device) { /* a later cfdp_configure() may name another device: nothing of this one survives */
    cudaStreamDestroy(E.s_comp); cudaStreamDestroy(E.s_comm); E.s_comp = E.s_comm = nullptr;
    cudaEventDestroy(E.ev_b); cudaEventDestroy(E.ev_x); cudaEventDestroy(E.ev_t0); cudaEventDestroy(E.ev_t1);
    E.ev_b = E.ev_x = E.ev_t0 = E.ev_t1 = nullptr;
    E.have_device = false;
  }

  E.d_var = E.d_grad = E.d_pvol = nullptr; E.d_blob = nullptr; E.d_tiles = nullptr; E.d_stage = nullptr;
  E.d_loc_dst = E.d_loc_src = E.d_send_rows = E.d_recv_rows = nullptr; E.d_sendbuf = E.d_recvbuf = nullptr;
  E.committed = false; E.planned = false; E.configured = false;
  E.h_exp_off.clear(); E.h_exp_src.clear(); E.h_exp_dst.clear(); E.d_exp_off = E.d_exp_src = E.d_exp_dst = nullptr;
  E.h_tiles.clear(); E.blob_base.clear(); E.h_loc_dst.clear(); E.h_loc_src.clear(); E.h_send_rows.clear(); E.h_recv_rows.clear();
  E.max_nfaces = E.max_nloc = E.max_npts = 0; E.max_stage = 0; E.blob_bytes = 0; E.max_blob = 0; E.max_nhalo = 0; E.max_footprint = 0; E.max_hvpv = 0; E.flux_smem = 0; E.with_flux = 0; E.flux_alg_bytes = 0; E.h_ftiles.clear(); E.fblob_bytes = 0;
  E.rows = E.ntiles = E.nbtiles = 0; E.nfaces = E.nown = E.nall = E.tile_faces = E.halo_refs = E.alg_bytes = 0;
  E.n_local = E.n_send = E.n_recv = 0; E.launches = 0; E.nprocs = 0; E.per_proc = 0;
}
