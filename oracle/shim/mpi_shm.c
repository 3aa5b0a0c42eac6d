/*
 * oracle/shim/mpi_shm.c -- TEST INFRASTRUCTURE, not product code.
 *
 * A tiny single-node MPI subset (fork + POSIX shared memory byte rings) that lets the
 * UNMODIFIED reference (/root/reference/src, one MPI rank per mesh domain) run in an
 * image without MPI.  Launched by mpirun_shim (same directory); without the launcher's
 * environment it degenerates to a single rank.
 *
 * Semantics implemented (all the reference needs):
 *   - MPI_THREAD_MULTIPLE (one process-wide mutex around the progress engine)
 *   - eager-free, in-order, per-(src,dst) matching of Isend/Irecv (tags are carried and
 *     checked but never used to reorder: the reference sends exactly one message per
 *     partner per phase -- src/comm_data.c:195-250, src/exchange_data_mpi.c:96-166)
 *   - Waitall / Waitany / Test / Testany, Send / Recv, Barrier
 *   - trivial Alloc_mem / Info / Win_allocate / Group calls (init_mpidma_buffers is called
 *     unconditionally, src/comm_data.c:500); the one-sided data path (Put/fence/PSCW)
 *     aborts -- it is only reachable with -DUSE_MPI_1_SIDED, which the rig does not set.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <stdatomic.h>
#include <pthread.h>
#include <sched.h>
#include <unistd.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include "mpi.h"

#define MAXREQ 4096
#define CACHELINE 128

typedef struct {
  _Atomic uint64_t head;               /* bytes produced */
  char pad0[CACHELINE - sizeof(uint64_t)];
  _Atomic uint64_t tail;               /* bytes consumed */
  char pad1[CACHELINE - sizeof(uint64_t)];
} ring_hdr;

typedef struct {
  _Atomic int bar_count;
  _Atomic int bar_sense;
  int size;
  int pad;
  uint64_t ring_bytes;
} shm_hdr;

typedef struct {
  int active;       /* slot in use */
  int kind;         /* 0 send, 1 recv */
  int peer;
  int tag;
  char *buf;
  size_t total;     /* payload bytes */
  size_t done;      /* payload bytes moved */
  int hdr_done;     /* header written / parsed */
  int complete;
  int next;         /* FIFO link */
  size_t got;       /* recv: actual message size */
} req_t;

static int g_rank = 0, g_size = 1;
static char *g_base = NULL;
static shm_hdr *g_hdr = NULL;
static uint64_t g_ring_bytes = 0;
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static req_t g_req[MAXREQ];
static int *sq_head, *sq_tail, *rq_head, *rq_tail; /* per peer FIFOs */
static int g_local_sense = 0;

static void die(const char *msg)
{
  fprintf(stderr, "[mpi_shm rank %d] fatal: %s\n", g_rank, msg);
  abort();
}

static ring_hdr *ring_of(int src, int dst)
{
  size_t stride = sizeof(ring_hdr) + g_ring_bytes;
  return (ring_hdr *)(g_base + 4096 + ((size_t)src * g_size + dst) * stride);
}

static size_t ring_put(ring_hdr *r, const char *src, size_t n)
{
  char *data = (char *)(r + 1);
  uint64_t head = atomic_load_explicit(&r->head, memory_order_relaxed);
  uint64_t tail = atomic_load_explicit(&r->tail, memory_order_acquire);
  size_t space = (size_t)(g_ring_bytes - (head - tail));
  if (n > space) n = space;
  if (!n) return 0;
  size_t off = (size_t)(head % g_ring_bytes);
  size_t first = n < g_ring_bytes - off ? n : g_ring_bytes - off;
  memcpy(data + off, src, first);
  if (n > first) memcpy(data, src + first, n - first);
  atomic_store_explicit(&r->head, head + n, memory_order_release);
  return n;
}

static size_t ring_space(ring_hdr *r)
{
  uint64_t head = atomic_load_explicit(&r->head, memory_order_relaxed);
  uint64_t tail = atomic_load_explicit(&r->tail, memory_order_acquire);
  return (size_t)(g_ring_bytes - (head - tail));
}

static size_t ring_avail(ring_hdr *r)
{
  uint64_t head = atomic_load_explicit(&r->head, memory_order_acquire);
  uint64_t tail = atomic_load_explicit(&r->tail, memory_order_relaxed);
  return (size_t)(head - tail);
}

static size_t ring_get(ring_hdr *r, char *dst, size_t n)
{
  char *data = (char *)(r + 1);
  uint64_t head = atomic_load_explicit(&r->head, memory_order_acquire);
  uint64_t tail = atomic_load_explicit(&r->tail, memory_order_relaxed);
  size_t avail = (size_t)(head - tail);
  if (n > avail) n = avail;
  if (!n) return 0;
  size_t off = (size_t)(tail % g_ring_bytes);
  size_t first = n < g_ring_bytes - off ? n : g_ring_bytes - off;
  memcpy(dst, data + off, first);
  if (n > first) memcpy(dst + first, data, n - first);
  atomic_store_explicit(&r->tail, tail + n, memory_order_release);
  return n;
}

/* ---- progress engine (call with g_lock held) ---- */
static void progress_peer(int p)
{
  /* sends to p */
  while (sq_head[p] >= 0) {
    req_t *q = &g_req[sq_head[p]];
    ring_hdr *r = ring_of(g_rank, p);
    if (!q->hdr_done) {
      if (ring_space(r) < 16) break;
      uint64_t h[2] = { (uint64_t)q->total, (uint64_t)(uint32_t)q->tag };
      ring_put(r, (const char *)h, 16);
      q->hdr_done = 1;
    }
    if (q->done < q->total) q->done += ring_put(r, q->buf + q->done, q->total - q->done);
    if (q->done < q->total) break;
    q->complete = 1;
    sq_head[p] = q->next;
    if (sq_head[p] < 0) sq_tail[p] = -1;
  }
  /* receives from p */
  while (rq_head[p] >= 0) {
    req_t *q = &g_req[rq_head[p]];
    ring_hdr *r = ring_of(p, g_rank);
    if (!q->hdr_done) {
      if (ring_avail(r) < 16) break;
      uint64_t h[2];
      ring_get(r, (char *)h, 16);
      q->got = (size_t)h[0];
      if ((int)(uint32_t)h[1] != q->tag) die("tag mismatch (in-order matching violated)");
      if (q->got > q->total) die("message longer than posted receive");
      q->hdr_done = 1;
    }
    if (q->done < q->got) q->done += ring_get(r, q->buf + q->done, q->got - q->done);
    if (q->done < q->got) break;
    q->complete = 1;
    rq_head[p] = q->next;
    if (rq_head[p] < 0) rq_tail[p] = -1;
  }
}

static void progress_all(void)
{
  for (int p = 0; p < g_size; p++)
    if (sq_head[p] >= 0 || rq_head[p] >= 0) progress_peer(p);
}

static int new_req(int kind, void *buf, size_t bytes, int peer, int tag)
{
  static int hint = 0;
  for (int n = 0; n < MAXREQ; n++) {
    int i = (hint + n) % MAXREQ;
    if (!g_req[i].active) {
      req_t *q = &g_req[i];
      memset(q, 0, sizeof *q);
      q->active = 1; q->kind = kind; q->peer = peer; q->tag = tag;
      q->buf = (char *)buf; q->total = bytes; q->next = -1;
      int *head = kind ? rq_head : sq_head, *tail = kind ? rq_tail : sq_tail;
      if (tail[peer] >= 0) g_req[tail[peer]].next = i; else head[peer] = i;
      tail[peer] = i;
      hint = i + 1;
      return i;
    }
  }
  die("request table exhausted");
  return -1;
}

static void relax(void)
{
  __asm__ __volatile__("rep; nop" ::: "memory");
}

/* ---- init / finalize ---- */
int MPI_Init_thread(int *argc, char ***argv, int required, int *provided)
{
  (void)argc; (void)argv; (void)required;
  if (provided) *provided = MPI_THREAD_MULTIPLE;
  const char *sz = getenv("CFDP_SHIM_SIZE"), *rk = getenv("CFDP_SHIM_RANK"), *nm = getenv("CFDP_SHIM_SHM");
  if (sz && rk && nm && atoi(sz) > 1) {
    g_size = atoi(sz);
    g_rank = atoi(rk);
    int fd = shm_open(nm, O_RDWR, 0600);
    if (fd < 0) die("shm_open failed");
    struct stat st;
    if (fstat(fd, &st)) die("fstat failed");
    g_base = mmap(NULL, (size_t)st.st_size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (g_base == MAP_FAILED) die("mmap failed");
    close(fd);
    g_hdr = (shm_hdr *)g_base;
    g_ring_bytes = g_hdr->ring_bytes;
    if (g_hdr->size != g_size) die("size mismatch");
  }
  sq_head = malloc(4 * g_size * sizeof(int));
  sq_tail = sq_head + g_size; rq_head = sq_tail + g_size; rq_tail = rq_head + g_size;
  for (int i = 0; i < 4 * g_size; i++) sq_head[i] = -1;
  return MPI_SUCCESS;
}

int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int *s) { (void)c; *s = g_size; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = g_rank; return MPI_SUCCESS; }

int MPI_Barrier(MPI_Comm c)
{
  (void)c;
  if (g_size == 1) return MPI_SUCCESS;
  pthread_mutex_lock(&g_lock);
  int sense = g_local_sense = !g_local_sense;
  if (atomic_fetch_add(&g_hdr->bar_count, 1) == g_size - 1) {
    atomic_store(&g_hdr->bar_count, 0);
    atomic_store(&g_hdr->bar_sense, sense);
  } else {
    while (atomic_load(&g_hdr->bar_sense) != sense) { progress_all(); relax(); sched_yield(); }
  }
  pthread_mutex_unlock(&g_lock);
  return MPI_SUCCESS;
}

/* ---- point to point ---- */
int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm c, MPI_Request *req)
{
  (void)dt; (void)c;
  if (dest < 0 || dest >= g_size || dest == g_rank) die("bad destination");
  pthread_mutex_lock(&g_lock);
  *req = new_req(0, (void *)buf, (size_t)count, dest, tag);
  progress_peer(dest);
  pthread_mutex_unlock(&g_lock);
  return MPI_SUCCESS;
}

int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm c, MPI_Request *req)
{
  (void)dt; (void)c;
  if (src < 0 || src >= g_size || src == g_rank) die("bad source");
  pthread_mutex_lock(&g_lock);
  *req = new_req(1, buf, (size_t)count, src, tag);
  progress_peer(src);
  pthread_mutex_unlock(&g_lock);
  return MPI_SUCCESS;
}

static void fill_status(MPI_Status *st, req_t *q)
{
  if (st) { st->MPI_SOURCE = q->peer; st->MPI_TAG = q->tag; st->MPI_ERROR = 0; st->count = (int)(q->kind ? q->got : q->total); }
}

/* returns 1 if request i was complete and has been retired */
static int retire_if_done(MPI_Request *r, MPI_Status *st)
{
  if (*r == MPI_REQUEST_NULL) return 1;
  if (*r < 0 || *r >= MAXREQ || !g_req[*r].active) die("invalid request handle");
  req_t *q = &g_req[*r];
  if (!q->complete) return 0;
  fill_status(st, q);
  q->active = 0;
  *r = MPI_REQUEST_NULL;
  return 1;
}

int MPI_Test(MPI_Request *req, int *flag, MPI_Status *st)
{
  pthread_mutex_lock(&g_lock);
  progress_all();
  *flag = retire_if_done(req, st);
  pthread_mutex_unlock(&g_lock);
  return MPI_SUCCESS;
}

static int wait_one(MPI_Request *req, MPI_Status *st)
{
  for (;;) {
    pthread_mutex_lock(&g_lock);
    progress_all();
    int done = retire_if_done(req, st);
    pthread_mutex_unlock(&g_lock);
    if (done) return MPI_SUCCESS;
    relax();
  }
}

int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *sts)
{
  for (int i = 0; i < n; i++) wait_one(&reqs[i], sts ? &sts[i] : NULL);
  return MPI_SUCCESS;
}

int MPI_Testany(int n, MPI_Request *reqs, int *index, int *flag, MPI_Status *st)
{
  pthread_mutex_lock(&g_lock);
  progress_all();
  int nactive = 0;
  *flag = 0; *index = MPI_UNDEFINED;
  for (int i = 0; i < n; i++) {
    if (reqs[i] == MPI_REQUEST_NULL) continue;
    nactive++;
    if (retire_if_done(&reqs[i], st)) { *flag = 1; *index = i; break; }
  }
  if (!nactive) *flag = 1;
  pthread_mutex_unlock(&g_lock);
  return MPI_SUCCESS;
}

int MPI_Waitany(int n, MPI_Request *reqs, int *index, MPI_Status *st)
{
  for (;;) {
    int flag;
    MPI_Testany(n, reqs, index, &flag, st);
    if (flag) return MPI_SUCCESS;
    relax();
  }
}

int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm c)
{
  MPI_Request r;
  MPI_Isend(buf, count, dt, dest, tag, c, &r);
  return wait_one(&r, NULL);
}

int MPI_Recv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm c, MPI_Status *st)
{
  MPI_Request r;
  MPI_Irecv(buf, count, dt, src, tag, c, &r);
  return wait_one(&r, st);
}

/* ---- memory / windows / groups: bookkeeping only ---- */
int MPI_Alloc_mem(MPI_Aint size, MPI_Info info, void *baseptr)
{
  (void)info;
  *(void **)baseptr = malloc(size > 0 ? (size_t)size : 1);
  return MPI_SUCCESS;
}
int MPI_Info_create(MPI_Info *info) { *info = 1; return MPI_SUCCESS; }
int MPI_Info_set(MPI_Info info, const char *k, const char *v) { (void)info; (void)k; (void)v; return MPI_SUCCESS; }
static int g_win_model = MPI_WIN_UNIFIED;
int MPI_Win_allocate(MPI_Aint size, int du, MPI_Info info, MPI_Comm c, void *baseptr, MPI_Win *win)
{
  (void)du; (void)info; (void)c;
  *(void **)baseptr = malloc(size > 0 ? (size_t)size : 1);
  *win = 1;
  return MPI_SUCCESS;
}
int MPI_Win_create(void *base, MPI_Aint size, int du, MPI_Info info, MPI_Comm c, MPI_Win *win)
{
  (void)base; (void)size; (void)du; (void)info; (void)c;
  *win = 1;
  return MPI_SUCCESS;
}
int MPI_Win_get_attr(MPI_Win win, int keyval, void *attr, int *flag)
{
  (void)win; (void)keyval;
  *(int **)attr = &g_win_model;
  *flag = 1;
  return MPI_SUCCESS;
}
int MPI_Win_free(MPI_Win *win) { *win = 0; return MPI_SUCCESS; }
int MPI_Comm_group(MPI_Comm c, MPI_Group *g) { (void)c; *g = 1; return MPI_SUCCESS; }
int MPI_Group_incl(MPI_Group g, int n, const int ranks[], MPI_Group *out) { (void)g; (void)n; (void)ranks; *out = 2; return MPI_SUCCESS; }
int MPI_Group_free(MPI_Group *g) { *g = 0; return MPI_SUCCESS; }

static int onesided(void) { die("one-sided MPI data path is not part of the oracle rig"); return 1; }
int MPI_Win_fence(int a, MPI_Win w) { (void)a; (void)w; return onesided(); }
int MPI_Win_post(MPI_Group g, int a, MPI_Win w) { (void)g; (void)a; (void)w; return onesided(); }
int MPI_Win_start(MPI_Group g, int a, MPI_Win w) { (void)g; (void)a; (void)w; return onesided(); }
int MPI_Win_complete(MPI_Win w) { (void)w; return onesided(); }
int MPI_Win_wait(MPI_Win w) { (void)w; return onesided(); }
int MPI_Put(const void *o, int oc, MPI_Datatype od, int t, MPI_Aint d, int tc, MPI_Datatype td, MPI_Win w)
{ (void)o; (void)oc; (void)od; (void)t; (void)d; (void)tc; (void)td; (void)w; return onesided(); }
