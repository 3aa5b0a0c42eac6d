/*
 * host_data.cpp -- host side of the drop-in boundary: the solver_data / comm_data containers,
 * the mesh loader calls and the halo tables, under the reference's own entry-point names.
 *
 *   read_solver_data / init_solver_data ........ reference src/solver_data.c:65-160
 *   init_communication ......................... reference src/comm_data.c:257-307 (MPI_Init_thread
 *        + rank/size become: rank/size of the torchrun-style environment, one GPU per process)
 *   read_communication_data .................... reference src/comm_data.c:74-114
 *   compute_communication_tables ............... reference src/comm_data.c:446-502 with
 *        create_recvsend_index (:116-255) and compute_offset_tables (:309-443)
 *   free_communication_ressources .............. reference src/comm_data.c:505-521
 * The reference maps one domain to one MPI rank and performs the sendindex handshake with
 * MPI_Send/MPI_Recv (tag 4711).  Here a process may host several domains: pairs hosted by the
 * same process are resolved by direct lookup, remote pairs through one grouped NCCL exchange.
 */
#include <string.h>
#include <algorithm>
#include <vector>
#include <omp.h>
#include "common.h"

static void *xmalloc(size_t bytes)
{
  ASSERT(bytes > 0); /* util.c:34: check_malloc(0) is an error in the reference as well */
  void *p = malloc(bytes);
  ASSERT(p != NULL);
  return p;
}

/* CFDP_LEAN_HOST=1 (device-resident runs on meshes of hundreds of millions of points): the host mirrors of the RESULTS
 * (sd->grad, sd->psd_flux: 192 bytes per point) are not allocated, and the mesh arrays sd->fpoint / sd->fnormal are
 * released once the GPU schedule has been built from them (cfdp_plan).  Entry points that need them fail loudly. */
static bool lean_host(void) { const char *e = getenv("CFDP_LEAN_HOST"); return e && atoi(e) != 0; }

/* ---------------------------------------------------------------------------------------- */
extern "C" void read_solver_data(int ncid, solver_data *sd)
{
  ASSERT(sd != NULL);
  memset(sd, 0, sizeof *sd);
  sd->ncolors = get_nc_val(ncid, "ncolors");
  sd->nfaces = get_nc_val(ncid, "nfaces");
  sd->nownpoints = get_nc_val(ncid, "nownpoints");
  sd->nallpoints = get_nc_val(ncid, "nallpoints");
  ASSERT(sd->ncolors > 0);
  ASSERT(sd->nfaces > 0);
  ASSERT(sd->nownpoints > 0);
  ASSERT(sd->nallpoints > 0);
  ASSERT(sd->nallpoints >= sd->nownpoints);
  const size_t nf = (size_t)sd->nfaces, na = (size_t)sd->nallpoints;
  sd->fpoint = (int(*)[2])xmalloc(nf * 2 * sizeof(int));
  sd->fnormal = (double(*)[3])xmalloc(nf * 3 * sizeof(double));
  sd->pvolume = (double *)xmalloc(na * sizeof(double));
  /* var / grad cross PCIe every drop-in call: page-locked when a CUDA device is present */
  sd->var = (double(*)[NGRAD])engine_alloc_pinned(na * NGRAD * sizeof(double));
  if (!lean_host()) {
    sd->grad = (double(*)[NGRAD][3])engine_alloc_pinned(na * NGRAD * 3 * sizeof(double));
    sd->psd_flux = (double(*)[NFLUX])xmalloc(na * NFLUX * sizeof(double));
  }
  get_nc_int(ncid, "fpoint", &sd->fpoint[0][0]);
  get_nc_double(ncid, "fnormal", &sd->fnormal[0][0]);
  get_nc_double(ncid, "pvolume", sd->pvolume);
  /* The file's colouring (fcolor_npoints / fcolor_points) is read and then discarded by the
   * reference (solver_data.c:126-158, threads.c:748-749); the GPU schedule does not use it. */
  sd->fcolor = NULL;
}

extern "C" void init_solver_data(solver_data *sd, int NITER)
{
  ASSERT(sd != NULL);
  ASSERT(sd->nallpoints != 0);
  const long long na = sd->nallpoints;
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < na; i++) {
    for (int j = 0; j < NGRAD; j++) sd->var[i][j] = 1.0;
    if (sd->grad) for (int j = 0; j < NGRAD; j++) for (int k = 0; k < 3; k++) sd->grad[i][j][k] = 1.0;
    if (sd->psd_flux) for (int j = 0; j < NFLUX; j++) sd->psd_flux[i][j] = 1.0;
  }
  sd->niter = NITER;
}

/* ---------------------------------------------------------------------------------------- */
static void zero_comm_data(comm_data *cd, int iProc, int nProc)
{
  memset(cd, 0, sizeof *cd); /* comm_data.c:35-71 */
  cd->nProc = nProc; cd->iProc = iProc;
}

static int env_int(const char *name, int dflt)
{
  const char *s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

extern "C" void init_communication(int argc, char *argv[], comm_data *cd)
{
  (void)argc; (void)argv;
  ASSERT(cd != NULL);
  Engine *e = engine_get();
  (void)e;
  cfdp_stats st; cfdp_get_stats(&st);
  if (st.nprocs == 0) {
    /* not configured: one domain per process, rank/size from the launcher's environment */
    const int rank = env_int("RANK", env_int("OMPI_COMM_WORLD_RANK", env_int("PMI_RANK", 0)));
    const int size = env_int("WORLD_SIZE", env_int("OMPI_COMM_WORLD_SIZE", env_int("PMI_SIZE", 1)));
    ASSERT(cfdp_configure(rank, size, size, -1) == 0);
    cfdp_get_stats(&st);
  }
  ASSERT(st.ndomains_hosted == 1); /* several domains per process: use cfdp_init_communication_domain */
  cfdp_init_communication_domain(cd, st.proc_rank);
}

extern "C" void cfdp_init_communication_domain(comm_data *cd, int domain)
{
  ASSERT(cd != NULL);
  cfdp_stats st; cfdp_get_stats(&st);
  ASSERT(st.nprocs > 0);
  const int ndom = st.nprocs * st.ndomains_hosted;
  ASSERT(domain >= 0 && domain < ndom);
  ASSERT(engine_proc_of_domain(domain) == st.proc_rank);
  zero_comm_data(cd, domain, ndom);
  engine_register_domain(cd, domain);
}

extern "C" void read_communication_data(int ncid, comm_data *cd)
{
  ASSERT(cd != NULL);
  Domain *d = engine_find_domain(cd);
  ASSERT(d != NULL); /* init_communication first */
  cd->ndomains = get_nc_val(ncid, "ndomains");
  cd->nownpoints = get_nc_val(ncid, "nownpoints");
  d->comm_read = true;
  if (cd->ndomains == 1) return;
  cd->naddpoints = get_nc_val(ncid, "naddpoints");
  cd->ncommdomains = get_nc_val(ncid, "ncommdomains");
  ASSERT(cd->ndomains >= 1);
  ASSERT(cd->ndomains == cd->nProc);
  ASSERT(cd->naddpoints > 0);
  ASSERT(cd->ncommdomains > 0);
  cd->commpartner = (int *)xmalloc((size_t)cd->ncommdomains * sizeof(int));
  cd->sendcount = (int *)xmalloc((size_t)cd->ndomains * sizeof(int));
  cd->recvcount = (int *)xmalloc((size_t)cd->ndomains * sizeof(int));
  cd->addpoint_owner = (int *)xmalloc((size_t)cd->naddpoints * sizeof(int));
  cd->addpoint_id = (int *)xmalloc((size_t)cd->naddpoints * sizeof(int));
  get_nc_int(ncid, "commpartner", cd->commpartner);
  get_nc_int(ncid, "sendcount", cd->sendcount);
  get_nc_int(ncid, "recvcount", cd->recvcount);
  get_nc_int(ncid, "addpoint_owner", cd->addpoint_owner);
  get_nc_int(ncid, "addpoint_idx", cd->addpoint_id);
}

static void attach_mesh(const cfdp_mesh_domain *m, comm_data *cd, solver_data *sd, cfdp_mesh_domain *take);
extern "C" void cfdp_attach_mesh(const cfdp_mesh_domain *m, comm_data *cd, solver_data *sd) { attach_mesh(m, cd, sd, NULL); }
/* the same, but the face and volume arrays (malloc'ed by the generator) change hands instead of being copied: at 8 M
 * points per domain the copies and their page faults were two thirds of the time it takes to generate a domain */
extern "C" void cfdp_attach_mesh_take(cfdp_mesh_domain *m, comm_data *cd, solver_data *sd) { attach_mesh(m, cd, sd, m); }

static void attach_mesh(const cfdp_mesh_domain *m, comm_data *cd, solver_data *sd, cfdp_mesh_domain *take)
{
  ASSERT(m != NULL && cd != NULL && sd != NULL);
  Domain *d = engine_find_domain(cd);
  ASSERT(d != NULL);
  memset(sd, 0, sizeof *sd);
  sd->ncolors = 1; sd->nfaces = m->nfaces; sd->nownpoints = m->nown; sd->nallpoints = m->nall;
  ASSERT(sd->nfaces > 0 && sd->nownpoints > 0);
  const size_t nf = (size_t)m->nfaces, na = (size_t)m->nall;
  if (take) {
    sd->fpoint = (int(*)[2])take->fpoint; sd->fnormal = (double(*)[3])take->fnormal; sd->pvolume = take->pvolume;
    take->fpoint = NULL; take->fnormal = NULL; take->pvolume = NULL;
  } else {
    sd->fpoint = (int(*)[2])xmalloc(nf * 2 * sizeof(int));
    sd->fnormal = (double(*)[3])xmalloc(nf * 3 * sizeof(double));
    sd->pvolume = (double *)xmalloc(na * sizeof(double));
    memcpy(sd->fpoint, m->fpoint, nf * 2 * sizeof(int));
    memcpy(sd->fnormal, m->fnormal, nf * 3 * sizeof(double));
    memcpy(sd->pvolume, m->pvolume, na * sizeof(double));
  }
  sd->var = (double(*)[NGRAD])engine_alloc_pinned(na * NGRAD * sizeof(double));
  if (!lean_host()) {
    sd->grad = (double(*)[NGRAD][3])engine_alloc_pinned(na * NGRAD * 3 * sizeof(double));
    sd->psd_flux = (double(*)[NFLUX])xmalloc(na * NFLUX * sizeof(double));
  }
  init_solver_data(sd, 25);
  cd->ndomains = m->ndomains; cd->nownpoints = m->nown;
  d->comm_read = true;
  if (cd->ndomains == 1) return;
  ASSERT(cd->ndomains == cd->nProc);
  cd->naddpoints = m->nadd; cd->ncommdomains = m->ncommdomains;
  ASSERT(cd->naddpoints > 0 && cd->ncommdomains > 0);
  cd->commpartner = (int *)xmalloc((size_t)m->ncommdomains * sizeof(int));
  cd->sendcount = (int *)xmalloc((size_t)m->ndomains * sizeof(int));
  cd->recvcount = (int *)xmalloc((size_t)m->ndomains * sizeof(int));
  cd->addpoint_owner = (int *)xmalloc((size_t)m->nadd * sizeof(int));
  cd->addpoint_id = (int *)xmalloc((size_t)m->nadd * sizeof(int));
  memcpy(cd->commpartner, m->commpartner, (size_t)m->ncommdomains * sizeof(int));
  memcpy(cd->sendcount, m->sendcount, (size_t)m->ndomains * sizeof(int));
  memcpy(cd->recvcount, m->recvcount, (size_t)m->ndomains * sizeof(int));
  memcpy(cd->addpoint_owner, m->addpoint_owner, (size_t)m->nadd * sizeof(int));
  memcpy(cd->addpoint_id, m->addpoint_idx, (size_t)m->nadd * sizeof(int));
}

/* ---------------------------------------------------------------------------------------- */
static void local_tables(comm_data *cd)
{
  const int nProc = cd->nProc, nown = cd->nownpoints, nadd = cd->naddpoints;
  ASSERT(cd->naddpoints != 0);
  ASSERT(cd->addpoint_owner != NULL);
  ASSERT(cd->addpoint_id != NULL);
  ASSERT(cd->commpartner != NULL);
  ASSERT(cd->sendcount != NULL);
  ASSERT(cd->recvcount != NULL);
  cd->sendindex = (int **)xmalloc((size_t)nProc * sizeof(int *));
  cd->recvindex = (int **)xmalloc((size_t)nProc * sizeof(int *));
  for (int i = 0; i < nProc; i++) { cd->sendindex[i] = NULL; cd->recvindex[i] = NULL; }
  for (int i = 0; i < cd->ncommdomains; i++) {
    const int k = cd->commpartner[i];
    ASSERT(k >= 0 && k < nProc && k != cd->iProc);
    if (cd->sendcount[k] > 0) {
      cd->sendindex[k] = (int *)xmalloc((size_t)cd->sendcount[k] * sizeof(int));
      for (int j = 0; j < cd->sendcount[k]; j++) cd->sendindex[k][j] = -1;
    }
    if (cd->recvcount[k] > 0) { /* comm_data.c:163-174 */
      int count = 0;
      cd->recvindex[k] = (int *)xmalloc((size_t)cd->recvcount[k] * sizeof(int));
      for (int j = 0; j < nadd; j++)
        if (cd->addpoint_owner[j] == k) { ASSERT(count < cd->recvcount[k]); cd->recvindex[k][count++] = nown + j; }
      ASSERT(count == cd->recvcount[k]);
    }
  }
  /* byte offsets of the per-partner regions in contiguous send / recv buffers (comm_data.c:343-352) */
  cd->local_recv_offset = (gaspi_offset_t *)xmalloc((size_t)nProc * sizeof(gaspi_offset_t));
  cd->local_send_offset = (gaspi_offset_t *)xmalloc((size_t)nProc * sizeof(gaspi_offset_t));
  cd->remote_recv_offset = (gaspi_offset_t *)xmalloc((size_t)nProc * sizeof(gaspi_offset_t));
  cd->notification = (gaspi_notification_id_t *)xmalloc((size_t)nProc * sizeof(gaspi_notification_id_t));
  for (int i = 0; i < nProc; i++) { cd->local_recv_offset[i] = cd->local_send_offset[i] = cd->remote_recv_offset[i] = 0; cd->notification[i] = 0; }
  gaspi_offset_t ssz = 0, rsz = 0;
  for (int i = 0; i < cd->ncommdomains; i++) {
    const int k = cd->commpartner[i];
    cd->local_send_offset[k] = ssz; cd->local_recv_offset[k] = rsz;
    ssz += (gaspi_offset_t)cd->sendcount[k] * CFDP_DIM2 * sizeof(double);
    rsz += (gaspi_offset_t)cd->recvcount[k] * CFDP_DIM2 * sizeof(double);
  }
  /* request / flag bookkeeping of init_mpi_requests (exchange_data_mpi.c:27-76); the staging
   * buffers themselves live on the device */
  cd->nreq = 2 * cd->ncommdomains;
  cd->req = NULL; cd->stat = NULL; cd->sendbuf = NULL; cd->recvbuf = NULL;
  cd->recv_flag = (volatile counter_t *)aligned_alloc(64, (size_t)cd->ncommdomains * sizeof(counter_t));
  cd->send_flag = (volatile counter_t *)aligned_alloc(64, (size_t)cd->ncommdomains * sizeof(counter_t));
  ASSERT(cd->recv_flag != NULL && cd->send_flag != NULL);
  for (int i = 0; i < cd->ncommdomains; i++) { cd->recv_flag[i].global = 0; cd->send_flag[i].global = 0; }
  cd->send_stage = cd->recv_stage = cd->comm_stage = 0;
}

static int slot_of(const comm_data *cd, int k)
{
  for (int i = 0; i < cd->ncommdomains; i++) if (cd->commpartner[i] == k) return i;
  return -1;
}

extern "C" void compute_communication_tables(comm_data *cd)
{
  ASSERT(cd != NULL);
  if (cd->ndomains == 1) return;
  Domain *me = engine_find_domain(cd);
  ASSERT(me != NULL);
  if (me->tables_done) return;
  /* collective over the hosted domains: all of them must have been read */
  const int nh = engine_num_hosted();
  for (int i = 0; i < nh; i++) { Domain *d = engine_hosted(i); ASSERT(d->cd != NULL && d->comm_read); }
  for (int i = 0; i < nh; i++) local_tables(engine_hosted(i)->cd);

  /* sendindex[k] = the partner's addpoint ids over ITS recvindex[me] order (comm_data.c:197-222) */
  struct Msg { int src, dst, proc; std::vector<int> buf; comm_data *cd; };
  std::vector<Msg> sends, recvs;
  for (int i = 0; i < nh; i++) {
    comm_data *a = engine_hosted(i)->cd;
    for (int s = 0; s < a->ncommdomains; s++) {
      const int k = a->commpartner[s];
      Domain *dk = engine_domain_by_id(k);
      if (dk) {
        comm_data *b = dk->cd;
        ASSERT(b->recvcount[a->iProc] == a->sendcount[k]);
        for (int j = 0; j < a->sendcount[k]; j++) a->sendindex[k][j] = b->addpoint_id[b->recvindex[a->iProc][j] - b->nownpoints];
        a->remote_recv_offset[k] = b->local_recv_offset[a->iProc]; /* comm_data.c:355-396 */
        const int sl = slot_of(b, a->iProc);
        ASSERT(sl >= 0);
        a->notification[k] = (gaspi_notification_id_t)sl;          /* comm_data.c:399-441 */
      } else {
        Msg ms; ms.src = a->iProc; ms.dst = k; ms.proc = engine_proc_of_domain(k); ms.cd = a;
        ms.buf.resize(3 + (size_t)a->recvcount[k]);
        const unsigned long long off = a->local_recv_offset[k];
        ms.buf[0] = (int)(off & 0xFFFFFFFFull); ms.buf[1] = (int)(off >> 32); ms.buf[2] = s;
        for (int j = 0; j < a->recvcount[k]; j++) ms.buf[3 + j] = a->addpoint_id[a->recvindex[k][j] - a->nownpoints];
        sends.push_back(std::move(ms));
        Msg mr; mr.src = k; mr.dst = a->iProc; mr.proc = engine_proc_of_domain(k); mr.cd = a;
        mr.buf.resize(3 + (size_t)a->sendcount[k]);
        recvs.push_back(std::move(mr));
      }
    }
  }
  if (!sends.empty() || !recvs.empty()) {
    auto key = [](const Msg &x, const Msg &y) { return x.src != y.src ? x.src < y.src : x.dst < y.dst; };
    std::sort(sends.begin(), sends.end(), key);
    std::sort(recvs.begin(), recvs.end(), key);
    std::vector<int> peer; std::vector<const int *> sb; std::vector<int> sc; std::vector<int *> rb; std::vector<int> rc;
    /* engine_exchange_ints issues sends and recvs per peer in list order: both sides sort by (src,dst) */
    for (auto &m : sends) { peer.push_back(m.proc); sb.push_back(m.buf.data()); sc.push_back((int)m.buf.size()); rb.push_back(nullptr); rc.push_back(0); }
    for (auto &m : recvs) { peer.push_back(m.proc); sb.push_back(nullptr); sc.push_back(0); rb.push_back(m.buf.data()); rc.push_back((int)m.buf.size()); }
    engine_exchange_ints(peer, sb, sc, rb, rc);
    for (auto &m : recvs) {
      comm_data *a = m.cd; const int k = m.src;
      a->remote_recv_offset[k] = (gaspi_offset_t)(unsigned)m.buf[0] | ((gaspi_offset_t)(unsigned)m.buf[1] << 32);
      a->notification[k] = (gaspi_notification_id_t)m.buf[2];
      for (int j = 0; j < a->sendcount[k]; j++) a->sendindex[k][j] = m.buf[3 + j];
    }
  }
  for (int i = 0; i < nh; i++) {
    Domain *d = engine_hosted(i);
    comm_data *a = d->cd;
    for (int s = 0; s < a->ncommdomains; s++) {
      const int k = a->commpartner[s];
      for (int j = 0; j < a->sendcount[k]; j++) ASSERT(a->sendindex[k][j] >= 0 && a->sendindex[k][j] < a->nownpoints);
    }
    d->tables_done = true;
  }
}

extern "C" void free_communication_ressources(comm_data *cd)
{
  ASSERT(cd != NULL);
  if (cd->ndomains == 1) return;
  /* the reference barriers, frees its MPI window and finalises MPI (comm_data.c:505-521) */
  cfdp_device_synchronize();
}
