/*
 * oracle/shim/mpirun_shim.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Launcher for the shared-memory MPI shim (mpi_shm.c):
 *     mpirun_shim -np N [-ring KiB] prog args...
 * creates one POSIX shm segment holding N*N byte rings, forks N ranks, waits for them and
 * kills the rest of the job when a rank fails.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <signal.h>
#include <unistd.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/wait.h>

typedef struct { int bar_count; int bar_sense; int size; int pad; uint64_t ring_bytes; } shm_hdr;

int main(int argc, char **argv)
{
  int np = 1, a = 1;
  uint64_t ring = 4u << 20;
  while (a < argc && argv[a][0] == '-') {
    if (!strcmp(argv[a], "-np") && a + 1 < argc) { np = atoi(argv[a + 1]); a += 2; }
    else if (!strcmp(argv[a], "-ring") && a + 1 < argc) { ring = (uint64_t)atol(argv[a + 1]) << 10; a += 2; }
    else break;
  }
  if (a >= argc || np < 1) { fprintf(stderr, "usage: %s -np N [-ring KiB] prog args...\n", argv[0]); return 2; }

  char name[64];
  snprintf(name, sizeof name, "/cfdp_shim_%d", (int)getpid());
  size_t total = 4096 + (size_t)np * np * (256 + ring);
  int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
  if (fd < 0) { perror("shm_open"); return 2; }
  if (ftruncate(fd, (off_t)total)) { perror("ftruncate"); shm_unlink(name); return 2; }
  shm_hdr *h = mmap(NULL, 4096, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  if (h == MAP_FAILED) { perror("mmap"); shm_unlink(name); return 2; }
  memset(h, 0, sizeof *h);
  h->size = np;
  h->ring_bytes = ring;
  close(fd);

  pid_t *pids = calloc((size_t)np, sizeof(pid_t));
  for (int r = 0; r < np; r++) {
    pid_t p = fork();
    if (p < 0) { perror("fork"); break; }
    if (p == 0) {
      char buf[32];
      snprintf(buf, sizeof buf, "%d", np); setenv("CFDP_SHIM_SIZE", buf, 1);
      snprintf(buf, sizeof buf, "%d", r);  setenv("CFDP_SHIM_RANK", buf, 1);
      setenv("CFDP_SHIM_SHM", name, 1);
      execvp(argv[a], &argv[a]);
      perror("execvp");
      _exit(127);
    }
    pids[r] = p;
  }
  int rc = 0, left = np;
  while (left > 0) {
    int st;
    pid_t p = wait(&st);
    if (p < 0) break;
    left--;
    int bad = !(WIFEXITED(st) && WEXITSTATUS(st) == 0);
    if (bad && !rc) {
      rc = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + WTERMSIG(st);
      for (int r = 0; r < np; r++) if (pids[r] != p && pids[r] > 0) kill(pids[r], SIGKILL);
    }
  }
  shm_unlink(name);
  return rc;
}
