/*
 * cfdp_b200.h -- C ABI of the B200-native Green-Gauss gradient + halo exchange path.
 *
 * Drop-in boundary for ONE hot path of CFD-Proxy: the Green-Gauss gradient face loop and
 * the halo exchange of `grad` that follows it.  The structs below are layout-compatible
 * with the reference's public structs and the entry points carry the reference's own
 * names, argument meaning and print-and-exit error convention, so that the reference
 * driver (src/hybrid.f6.c:27-101) can be linked against libcfdp_b200.so instead of its
 * own gradients.c / rangelist.c / threads.c / exchange_data_*.c / comm_data.c /
 * solver_data.c / read_netcdf.c.  Everything is extern "C", plain pointers and sizes.
 *
 * Reference interfaces replaced (file:line in /root/reference/src):
 *   solver_data, RangeList ........ solver_data.h:28-81
 *   comm_data ..................... comm_data.h:15-55
 *   init_communication ............ comm_data.h:58   (comm_data.c:257-307)
 *   read_communication_data ....... comm_data.h:59   (comm_data.c:74-114)
 *   compute_communication_tables .. comm_data.h:60   (comm_data.c:446-502, :116-255)
 *   free_communication_ressources . comm_data.h:61   (comm_data.c:505-521)
 *   read_solver_data .............. solver_data.h:85 (solver_data.c:80-160)
 *   init_solver_data .............. solver_data.h:84 (solver_data.c:65-77)
 *   init_threads .................. rangelist.h:17-20 (threads.c:730-788)  -> builds the GPU face schedule
 *   compute_gradients_gg_* ........ gradients.h:7-25 (gradients.c:150-335)
 *   exchange_dbl_mpi_post_recv .... exchange_data_mpi.h:37 (exchange_data_mpi.c:134-166)
 *   get_nc_val/get_nc_int/get_nc_double  read_netcdf.h:4-6 (read_netcdf.c:20-61)
 *   nc_open, nc_close, nc_inq_X, nc_get_var_X (libnetcdf subset used by the reference) -> cfdp_nc_X
 */
#ifndef CFDP_B200_H
#define CFDP_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGRAD 7
#define NFLUX 3
#define CFDP_DIM2 21 /* NGRAD*3 doubles per halo row (gradients.c:155 passes dim2 = NGRAD*3) */

/* ------------------------------------------------------------------------------------------
 * Reference-layout structs (field order, types and names as in the reference headers).
 * MPI handle arrays are kept as opaque pointers: this library has no MPI.
 * ---------------------------------------------------------------------------------------- */
typedef struct { int global __attribute__((aligned(64))); } counter_t; /* solver_data.h:12-15 */

typedef struct RangeList_t { /* solver_data.h:28-63; only kept for layout compatibility */
  struct RangeList_t *succ;
  int start, stop, ftype;
  int nall_points_of_color;   int *all_points_of_color;
  int nfirst_points_of_color; int *first_points_of_color;
  int nlast_points_of_color;  int *last_points_of_color;
  int nsendcount; int *sendpartner; int *sendcount; int **sendindex; int **sendoffset;
  int nrecvcount; int *recvpartner; int *recvcount; int **recvindex; int **recvoffset;
  int tid;
} RangeList;

typedef struct { /* solver_data.h:66-81 */
  int nfaces;
  int nallfaces;
  int nownpoints;
  int nallpoints;
  int ncolors;
  int (*fpoint)[2];
  double (*fnormal)[3];
  double *pvolume;
  double (*var)[NGRAD];
  double (*grad)[NGRAD][3];
  double (*psd_flux)[NFLUX];
  RangeList *fcolor;
  int niter;
} solver_data;

typedef unsigned long gaspi_offset_t;          /* comm_data.h:9-10 (non-GASPI build) */
typedef unsigned short gaspi_notification_id_t;

typedef struct { /* comm_data.h:15-55 */
  int nProc;
  int iProc;
  int ndomains;
  int ncommdomains;
  int nownpoints;
  int naddpoints;
  int *addpoint_owner;
  int *addpoint_id;
  int *commpartner;
  int *sendcount;   /* [ndomains], rank-indexed */
  int *recvcount;   /* [ndomains], rank-indexed */
  int **recvindex;  /* [nProc][recvcount[k]] local ghost ids = nown + j            (comm_data.c:163-174) */
  int **sendindex;  /* [nProc][sendcount[k]] local own ids partner k reads from us (comm_data.c:197-222) */
  int nreq;
  void *req;        /* MPI_Request* in the reference; unused here */
  void *stat;       /* MPI_Status*  in the reference; unused here */
  double **recvbuf; /* [ncommdomains] slot-indexed; device staging lives inside the library */
  double **sendbuf;
  gaspi_offset_t *remote_recv_offset;
  gaspi_offset_t *local_recv_offset;
  gaspi_offset_t *local_send_offset;
  gaspi_notification_id_t *notification;
  volatile counter_t *recv_flag;
  volatile counter_t *send_flag;
  volatile int recv_stage;
  volatile int send_stage;
  volatile int comm_stage;
} comm_data;

/* ------------------------------------------------------------------------------------------
 * Reference-named entry points (same argument meaning; errors print
 * "Error: '<expr>' [file:line]" to stderr and exit(EXIT_FAILURE), error_handling.h:29-36).
 * ---------------------------------------------------------------------------------------- */
void init_communication(int argc, char *argv[], comm_data *cd);
void read_communication_data(int ncid, comm_data *cd);
void compute_communication_tables(comm_data *cd);
void free_communication_ressources(comm_data *cd);
void read_solver_data(int ncid, solver_data *sd);
void init_solver_data(solver_data *sd, int NITER);
void init_threads(comm_data *cd, solver_data *sd, int NTHREADS);

void compute_gradients_gg_comm_free(comm_data *cd, solver_data *sd, int final);
void compute_gradients_gg_mpi_bulk_sync(comm_data *cd, solver_data *sd, int final);
void compute_gradients_gg_mpi_early_recv(comm_data *cd, solver_data *sd, int final);
void compute_gradients_gg_mpi_async(comm_data *cd, solver_data *sd, int final);
void compute_gradients_gg_gaspi_bulk_sync(comm_data *cd, solver_data *sd, int final);
void compute_gradients_gg_gaspi_async(comm_data *cd, solver_data *sd, int final);
void compute_gradients_gg_mpifence_bulk_sync(comm_data *cd, solver_data *sd, int final);
void compute_gradients_gg_mpifence_async(comm_data *cd, solver_data *sd, int final);
void compute_gradients_gg_mpipscw_bulk_sync(comm_data *cd, solver_data *sd, int final);
void compute_gradients_gg_mpipscw_async(comm_data *cd, solver_data *sd, int final);
void exchange_dbl_mpi_post_recv(comm_data *cd, int dim2);
/* flux.h:12 (flux.c:193-201), the consumer of the exchanged gradients, called once per iteration after
 * compute_gradients_gg_* (solver.c:52).  Reads sd->grad[p][IVX..IVZ][0..2] of both end points of every face,
 * writes the own rows of sd->psd_flux.  Results are those of the reference run with one OpenMP thread (with more
 * threads the reference's `ftype` tests make the result depend on the thread partition, flux.c:179-190). */
void compute_psd_flux(solver_data *sd);

int  get_nc_val(int ncid, const char *name);
void get_nc_int(int ncid, const char *name, int *array);
void get_nc_double(int ncid, const char *name, double *array);

/* NetCDF-3 classic (CDF-1/CDF-2) reader: the libnetcdf subset the reference calls
 * (hybrid.f6.c:65-66,89-90; read_netcdf.c:24-58).  include/netcdf_compat.h maps nc_* to these. */
#define CFDP_NC_NOWRITE 0
int cfdp_nc_open(const char *path, int mode, int *ncidp);
int cfdp_nc_close(int ncid);
const char *cfdp_nc_strerror(int err);
int cfdp_nc_inq_dimid(int ncid, const char *name, int *dimidp);
int cfdp_nc_inq_dimlen(int ncid, int dimid, size_t *lenp);
int cfdp_nc_inq_varid(int ncid, const char *name, int *varidp);
int cfdp_nc_get_var_int(int ncid, int varid, int *ip);
int cfdp_nc_get_var_double(int ncid, int varid, double *dp);

/* ------------------------------------------------------------------------------------------
 * Extensions (cfdp_ prefix).  The reference maps one mesh domain to one MPI rank; here a
 * process drives ONE GPU and may host several domains ("virtual ranks"): domains
 * [first, first+count) with count = ndomains/nprocs, first = proc_rank*count.
 * ---------------------------------------------------------------------------------------- */
enum { CFDP_COMM_FREE = 0, CFDP_MPI_BULK_SYNC = 1, CFDP_MPI_EARLY_RECV = 2, CFDP_MPI_ASYNC = 3,
       CFDP_GASPI_BULK_SYNC = 4, CFDP_GASPI_ASYNC = 5 };

/* process/GPU placement; call before init_communication.  device < 0: use LOCAL_RANK (or 0). */
int  cfdp_configure(int proc_rank, int nprocs, int ndomains_total, int device);
/* like init_communication but for the hosted domain `domain` (cd->iProc = domain, cd->nProc = ndomains) */
void cfdp_init_communication_domain(comm_data *cd, int domain);
/* NCCL bootstrap for nprocs > 1: rank 0 creates the 128-byte id, the caller ships it (any transport) */
int  cfdp_nccl_get_unique_id(void *id128);
int  cfdp_nccl_init(const void *id128);
/* transport of the setup-time index handshake between processes (comm_data.c:195-250: MPI_Send/MPI_Recv
 * tag 4711).  Default: NCCL.  Messages i = 0..n-1 in list order; per peer the i-th send matches the
 * peer's i-th receive.  scount[i] ints from sbuf[i] go to process peer[i], rcount[i] ints arrive in rbuf[i]. */
typedef void (*cfdp_int_exchange_fn)(int n, const int *peer, const int *const *sbuf, const int *scount,
                                     int *const *rbuf, const int *rcount);
void cfdp_set_int_exchange(cfdp_int_exchange_fn fn);
/* the per-peer plan of the halo exchange: for peer process index i (0..npeers-1) returns the peer's process
 * rank and the numbers of grad rows sent / received per iteration; rows (device rows) may be NULL */
int cfdp_get_peer_plan(int i, int *proc, long long *send_rows, long long *recv_rows);
/* host point (domain rank, local point id) behind entry j of the packed send (dir = 0) / recv (dir = 1) buffer */
int cfdp_get_exchange_entry(int dir, long long j, int *domain, int *point);
/* export list of boundary tile `tile` (index in the per-GPU tile list, < nboundary_tiles): the rows the gradient kernel
 * writes for other domains while it still holds them (fused pack).  Entry i: src_row[i] = device row inside the tile,
 * kind[i] = 0: dst[i] is the device row of a ghost point of a domain hosted on this GPU; kind[i] = 1: dst[i] is a slot
 * of the packed send buffer (see cfdp_get_exchange_entry(0, slot, ...)).  Returns the number of entries (arrays may
 * be NULL), -1 on error.  The analogue of the per-colour send lists checked by thread_comm.c:159-205. */
int cfdp_get_tile_exports(int tile, int capacity, unsigned *src_row, unsigned *dst, int *kind);
/* device row -> (hosted domain rank, local point id); -1 when the row is alignment padding */
int cfdp_get_row_owner(long long row, int *domain, int *point);
/* called once after init_threads() of every hosted domain (implicit on first compute call):
 * builds the unified device layout, the pack/unpack lists and the exchange plan */
void cfdp_commit(void);
/* host half of cfdp_commit (face schedules, device row numbering, pack/unpack row lists); needs no GPU */
void cfdp_plan(void);
/* host<->device mirrors (SURVEY 8(b) ownership): var is uploaded, grad downloaded */
void cfdp_var_to_device(solver_data *sd);
void cfdp_grad_to_host(solver_data *sd);
void cfdp_grad_to_device(solver_data *sd);   /* all rows of sd->grad, ghosts included */
void cfdp_flux_to_host(solver_data *sd);     /* own rows of sd->psd_flux */
/* resident = 1: compute_gradients_gg_* leave var/grad on the device (no per-call PCIe copies);
 * resident = 0 (default): every call uploads sd->var and downloads sd->grad (true drop-in) */
void cfdp_set_resident(int resident);
/* exact = 1 (default): separate multiply and add in the reference's single-thread summation order
 * (bit-identical to the reference run with 1 thread); exact = 0: fused multiply-add */
void cfdp_set_exact(int exact);
/* after cfdp_commit: which gradient kernel runs and how its grid walks the tile list.  version 2 = production
 * (gg_tile_pipe_kernel), 1 = one tile per CTA (second, independent implementation used by the tests to cross-check);
 * chunk = consecutive tiles per CTA; persistent > 0 = that many
 * CTAs walk all tiles, interleaved.  Returns the version in effect, -1 on error.  Defaults: CFDP_KERNEL / CFDP_CHUNK /
 * CFDP_PERSISTENT or 2 / 8 / 0. */
int cfdp_set_kernel(int version, int chunk, int persistent);
/* debug (CFDP_PHASE_PROF=1): SM cycles of thread 0 summed over tiles: [0] wait for data, [1] face walk, [2] rest, [4] tiles */
int cfdp_get_phase_profile(unsigned long long *out8, int reset);
/* run `niter` iterations of variant over ALL hosted domains, device resident; returns the
 * device time in milliseconds (CUDA events on the compute stream) */
double cfdp_iterate(int variant, int niter, int final_last);
/* A solver that keeps var on the device and changes it there (the reference only ever reads sd->var, solver.c:45-55) tells
 * the library so: the derived per-tile copies of the halo var rows are rebuilt from the device var rows.
 * cfdp_refresh_var(niter) does that niter times and returns the device time in ms; cfdp_set_var_refresh(1) makes every
 * iteration of cfdp_iterate start with it (the cost model "var is new in every iteration", bench.py `var_refresh`). */
double cfdp_refresh_var(int niter);
void cfdp_set_var_refresh(int on);
/* on = 1: every iteration of cfdp_iterate also runs the pseudo flux after the exchange (solver.c:45-55) */
void cfdp_set_flux(int on);
/* `niter` pseudo-flux passes over all hosted domains on the device grad as it stands; device time in ms */
double cfdp_flux_iterate(int niter);
/* one end-to-end step over all hosted domains: H2D var, iterate once, D2H grad (host buffers) */
double cfdp_step_e2e(int variant);
void cfdp_device_synchronize(void);
void cfdp_finalize(void);

typedef struct {
  long long nfaces;          /* faces computed per iteration on this process (sum over hosted domains) */
  long long nown, nall;      /* points (sum over hosted domains) */
  long long rows;            /* device rows incl. alignment padding */
  long long ntiles, nboundary_tiles;
  long long tile_faces;      /* face records stored in tile blobs (cut faces duplicated) */
  long long halo_refs;       /* var rows gathered from outside tiles per iteration */
  long long blob_bytes;      /* bytes of static schedule data read per iteration */
  long long send_rows_local, send_rows_remote; /* halo rows copied on-device / shipped over NCCL */
  long long alg_bytes;       /* F*32 + P_all*56 + P_own*176 */
  long long h2d_bytes, d2h_bytes; /* per e2e step */
  long long launches;        /* kernels launched by this library so far */
  long long lds_wavefronts_min, lds_wavefronts_est; /* schedule quality: shared-memory wavefronts of the face walk, conflict-free vs estimated */
  double last_kernel_ms;     /* mean device time of the gradient kernel(s) per iteration in the last cfdp_iterate */
  int nprocs, proc_rank, ndomains_hosted, tile_points, smem_bytes;
  int flux_smem_bytes;       /* shared memory per CTA of the pseudo-flux kernel */
  long long flux_alg_bytes;  /* algorithmic bytes of one pseudo-flux pass: 32 B per face + 72 B per point + 24 B per own point */
  double last_flux_ms;       /* mean device time of one pseudo-flux pass in the last cfdp_flux_iterate */
  long long flux_blob_bytes; /* size of the pseudo-flux tile blobs (0: the kernel reads the gradient blobs) */
  long long halo_pack_bytes; /* packed halo rows (one contiguous block per tile, refreshed at every upload of var): bytes the gradient kernel reads on top of the algorithmic ones */
  long long device_bytes;    /* device memory this library holds */
  int transport;             /* what the last exchange used: 0 none, 1 on-device copies only, 2 NCCL send/recv, 3 CUDA-IPC put + notify, 4 direct stores into peer memory */
  int ipc_ready, direct_ready, loopback;
} cfdp_stats;
void cfdp_get_stats(cfdp_stats *st);

/* schedule introspection for tests (invariants of eval.c:88-235 restated for the GPU schedule) */
typedef struct {
  int ntiles, nboundary_tiles, nrows;   /* rows of this domain incl. padding and ghosts */
  const int *row_of_point;              /* [nallpoints] device row (relative to the domain's first row) */
  const int *tile_row0;                 /* [ntiles+1] */
  const int *tile_npts;                 /* [ntiles] */
  const int *tile_nfaces;               /* [ntiles] */
  const int *tile_nhalo;                /* [ntiles] */
  const int *tile_is_boundary;          /* [ntiles] */
} cfdp_schedule_view;
int cfdp_get_schedule(const solver_data *sd, cfdp_schedule_view *v);
/* raw tile contents for tests: returns counts, fills caller buffers when non-NULL */
int cfdp_get_tile(const solver_data *sd, int tile, int *face_ids /*[nfaces]*/, int *halo_points /*[nhalo]*/);
/* raw tile blob (host copy, between cfdp_plan and cfdp_commit): which = 0 gradient blob, 1 pseudo-flux blob;
 * desc8 = {row0, npts, nhalo, nfaces | zslot << 16, maxdeg, npad, blob_bytes, halo_off}; returns blob_bytes or -1.  Layout:
 * [normals nfaces*3 f64][pad16][halo device rows nhalo u32][pad16][ELL maxdeg*npad u32]; ELL entry = tile-local
 * point | ghost << 15 | face slot << 16 | (point is p1) << 31, 0xFFFFFFFF = none; local >= CFDP_HALO_BASE(npts) = halo */
long long cfdp_get_tile_blob(const solver_data *sd, int tile, int which, unsigned *desc8, unsigned char *bytes, long long capacity);
/* device-side halo lists of a hosted domain in HOST numbering: rows packed for partner k and rows unpacked from k */
int cfdp_get_pack_list(const comm_data *cd, int partner, int *points /*[sendcount[partner]]*/);
int cfdp_get_unpack_list(const comm_data *cd, int partner, int *points /*[recvcount[partner]]*/);
/* copy of the last packed send rows for `partner` (sendcount*21 doubles) -- parity hook for threads.c:791-813 */
int cfdp_get_sendbuf(const comm_data *cd, int partner, double *rows);

/* ------------------------------------------------------------------------------------------
 * Synthetic F6-schema meshes (csrc/mesh_gen.c)
 * ---------------------------------------------------------------------------------------- */
enum { CFDP_ORDER_LEX = 0, CFDP_ORDER_BRICK = 1, CFDP_ORDER_SHUFFLE = 2 };
typedef struct {
  int nx, ny, nz;        /* global lattice */
  int px, py, pz;        /* domain grid: ndomains = px*py*pz */
  int order;             /* numbering of own points inside a domain */
  int brick;             /* brick edge for CFDP_ORDER_BRICK */
  int hexcut;            /* diagonal edges exist only for lower endpoints with x >= hexcut */
  int allow_big;         /* lift the reference's int-overflow size limits (GPU-only meshes) */
  double jitter;         /* relative perturbation of the face normals */
  unsigned long long seed;
} cfdp_mesh_spec;
typedef struct {
  int nfaces, nown, nall, nadd, ndomains, ncommdomains;
  int *fpoint;           /* [nfaces][2] */
  double *fnormal;       /* [nfaces][3] */
  double *pvolume;       /* [nall] */
  int *commpartner;      /* [ncommdomains] */
  int *sendcount, *recvcount;          /* [ndomains] */
  int *addpoint_owner, *addpoint_idx;  /* [nadd] */
  long long *global_id;  /* [nall] lattice id, seeds var */
} cfdp_mesh_domain;
int  cfdp_mesh_num_domains(const cfdp_mesh_spec *s);
long long cfdp_mesh_count_faces_global(const cfdp_mesh_spec *s);
int  cfdp_mesh_gen_domain(const cfdp_mesh_spec *s, int rank, cfdp_mesh_domain *out);
void cfdp_mesh_free_domain(cfdp_mesh_domain *m);
void cfdp_mesh_fill_var(const cfdp_mesh_domain *m, unsigned long long seed, double *var);
double cfdp_mesh_var_value(unsigned long long seed, long long gid, int eq);
/* fill sd/cd from an in-memory domain exactly as read_solver_data + init_solver_data +
 * read_communication_data would from its NetCDF file (arrays are copied) */
void cfdp_attach_mesh(const cfdp_mesh_domain *m, comm_data *cd, solver_data *sd);
/* the same without copies: the face and volume arrays of *m change hands (m keeps the rest; free it as usual) */
void cfdp_attach_mesh_take(cfdp_mesh_domain *m, comm_data *cd, solver_data *sd);

#ifdef __cplusplus
}
#endif
#endif
