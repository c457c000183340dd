// pcie_bw.cu -- what the BOX gives the end-to-end path: N processes, one per GPU, each
// copying pinned host memory to its GPU and back at the same time, with the chunk sizes
// bench.py's staged import uses (16 frames of 1080p: 132.7 MB in, 82.9 MB out).
//
//   nvcc -O2 -o pcie_bw pcie_bw.cu
//   ./pcie_bw [--gpus 1,2,4,8] [--seconds 0.6] [--quick]
//
// One JSON line per (gpu count, direction mix, variant).  Variants: default pinned memory;
// write-combined source for the H2D side; process pinned to the CPUs of the GPU's NUMA node
// before the pinned buffers are first touched; 32 MB chunks; two streams per direction.
// The children are forked BEFORE any CUDA call; they meet at barriers kept in a shared
// anonymous mapping, run each case for a fixed wall time and count the bytes whose copies
// completed.  Also reports a plain host memcpy bandwidth (T threads), the figure a
// host-DRAM-bound box would show.
#include <cuda_runtime.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <string>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); _exit(3); } } while (0)

struct Shared {
  std::atomic<int> arrive[64];
  double h2d_bytes[16], d2h_bytes[16], secs[16];
  int numa_node[16];
};

static double now() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void barrier(Shared* sh, int slot, int n) {
  sh->arrive[slot].fetch_add(1);
  while (sh->arrive[slot].load() < n) usleep(50);
}

struct Case { const char* name; int h2d, d2h; };
struct Variant { const char* name; bool wc, affine; size_t in_chunk, out_chunk; int streams; };

static int numa_node_of(int dev) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) != cudaSuccess) return -1;
  for (char* p = bus; *p; ++p) if (*p >= 'A' && *p <= 'F') *p += 32;
  char path[128];
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE* f = fopen(path, "r");
  if (!f) return -1;
  int n = -1;
  if (fscanf(f, "%d", &n) != 1) n = -1;
  fclose(f);
  return n;
}

static bool bind_to_node(int node) {
  if (node < 0) return false;
  char path[128];
  snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
  FILE* f = fopen(path, "r");
  if (!f) return false;
  char buf[4096] = {0};
  if (!fgets(buf, sizeof(buf), f)) { fclose(f); return false; }
  fclose(f);
  cpu_set_t set;
  CPU_ZERO(&set);
  for (char* tok = strtok(buf, ",\n"); tok; tok = strtok(NULL, ",\n")) {
    int a, b;
    if (sscanf(tok, "%d-%d", &a, &b) == 2) { for (int c = a; c <= b; ++c) CPU_SET(c, &set); }
    else if (sscanf(tok, "%d", &a) == 1) CPU_SET(a, &set);
  }
  return sched_setaffinity(0, sizeof(set), &set) == 0;
}

static void child(int g, int n, Shared* sh, const std::vector<Case>& cases, const std::vector<Variant>& vars,
                  double seconds, int slot0) {
  CK(cudaSetDevice(g));
  sh->numa_node[g] = numa_node_of(g);
  const size_t total = (size_t)1 << 30;
  uint8_t *d_in, *d_out;
  CK(cudaMalloc(&d_in, total));
  CK(cudaMalloc(&d_out, total));
  CK(cudaMemset(d_out, 1, total));
  int slot = slot0;
  cpu_set_t all;
  sched_getaffinity(0, sizeof(all), &all);
  for (const Variant& v : vars) {
    if (v.affine) bind_to_node(sh->numa_node[g]); else sched_setaffinity(0, sizeof(all), &all);
    uint8_t *h_in, *h_out;
    CK(cudaHostAlloc((void**)&h_in, total, v.wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
    CK(cudaHostAlloc((void**)&h_out, total, cudaHostAllocDefault));
    memset(h_in, 3, total);   // first touch here, under the affinity of this variant
    memset(h_out, 0, total);
    cudaStream_t si[4], so[4];
    for (int i = 0; i < v.streams; ++i) { CK(cudaStreamCreateWithFlags(&si[i], cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&so[i], cudaStreamNonBlocking)); }
    for (const Case& c : cases) {
      barrier(sh, slot++, n);
      const double t0 = now();
      double hb = 0, db = 0;
      size_t oi = 0, oo = 0;
      int k = 0;
      // keep two chunks per stream in flight: enqueue, then wait for the older one
      cudaEvent_t evi[8], evo[8];
      for (int i = 0; i < 8; ++i) { CK(cudaEventCreateWithFlags(&evi[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&evo[i], cudaEventDisableTiming)); }
      const int depth = 2 * v.streams;
      while (now() - t0 < seconds) {
        const int e = k % depth;
        if (k >= depth) {
          if (c.h2d) { CK(cudaEventSynchronize(evi[e])); hb += v.in_chunk; }
          if (c.d2h) { CK(cudaEventSynchronize(evo[e])); db += v.out_chunk; }
        }
        if (c.h2d) {
          if (oi + v.in_chunk > total) oi = 0;
          CK(cudaMemcpyAsync(d_in + oi, h_in + oi, v.in_chunk, cudaMemcpyHostToDevice, si[k % v.streams]));
          CK(cudaEventRecord(evi[e], si[k % v.streams]));
          oi += v.in_chunk;
        }
        if (c.d2h) {
          if (oo + v.out_chunk > total) oo = 0;
          CK(cudaMemcpyAsync(h_out + oo, d_out + oo, v.out_chunk, cudaMemcpyDeviceToHost, so[k % v.streams]));
          CK(cudaEventRecord(evo[e], so[k % v.streams]));
          oo += v.out_chunk;
        }
        ++k;
      }
      const double t1 = now();   // bytes completed by now; the tail in flight is not counted
      CK(cudaDeviceSynchronize());
      for (int i = 0; i < 8; ++i) { cudaEventDestroy(evi[i]); cudaEventDestroy(evo[i]); }
      sh->h2d_bytes[g] = hb; sh->d2h_bytes[g] = db; sh->secs[g] = t1 - t0;
      barrier(sh, slot++, n);
      barrier(sh, slot++, n);   // parent has printed
    }
    for (int i = 0; i < v.streams; ++i) { cudaStreamDestroy(si[i]); cudaStreamDestroy(so[i]); }
    CK(cudaFreeHost(h_in));
    CK(cudaFreeHost(h_out));
  }
  _exit(0);
}

struct MemArg { uint8_t *a, *b; size_t n; double secs; double bytes; };
static void* memcpy_thread(void* p) {
  MemArg* m = (MemArg*)p;
  const double t0 = now();
  double bytes = 0;
  while (now() - t0 < m->secs) { memcpy(m->b, m->a, m->n); bytes += m->n; memcpy(m->a, m->b, m->n); bytes += m->n; }
  m->bytes = bytes / (now() - t0);
  return NULL;
}

static void host_memcpy_bw(int threads, double secs) {
  std::vector<MemArg> args(threads);
  std::vector<pthread_t> th(threads);
  const size_t n = (size_t)128 << 20;
  for (int i = 0; i < threads; ++i) {
    args[i].a = (uint8_t*)malloc(n); args[i].b = (uint8_t*)malloc(n); args[i].n = n; args[i].secs = secs;
    memset(args[i].a, 1, n); memset(args[i].b, 2, n);
  }
  for (int i = 0; i < threads; ++i) pthread_create(&th[i], NULL, memcpy_thread, &args[i]);
  double total = 0;
  for (int i = 0; i < threads; ++i) { pthread_join(th[i], NULL); total += args[i].bytes; free(args[i].a); free(args[i].b); }
  printf("{\"host_memcpy\": {\"threads\": %d, \"copied_gbs\": %.1f, \"dram_traffic_gbs\": %.1f}}\n", threads, total / 1e9, 2 * total / 1e9);
  fflush(stdout);
}

int main(int argc, char** argv) {
  std::vector<int> counts;
  double seconds = 0.6;
  bool quick = false;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--gpus") && i + 1 < argc) {
      char* s = argv[++i];
      for (char* t = strtok(s, ","); t; t = strtok(NULL, ",")) counts.push_back(atoi(t));
    } else if (!strcmp(argv[i], "--seconds") && i + 1 < argc) seconds = atof(argv[++i]);
    else if (!strcmp(argv[i], "--quick")) quick = true;
  }
  // device count from a throw-away child so that the parent never initialises CUDA
  int pfd[2];
  if (pipe(pfd)) return 2;
  if (fork() == 0) { int n = 0; cudaGetDeviceCount(&n); if (write(pfd[1], &n, sizeof(n)) < 0) {} _exit(0); }
  int have = 0;
  if (read(pfd[0], &have, sizeof(have)) < 0) have = 0;
  wait(NULL);
  if (counts.empty()) for (int n = 1; n <= have; n *= 2) counts.push_back(n);
  const size_t IN = (size_t)16 * 1920 * 1080 * 4, OUT = (size_t)16 * (1920 * 1080 * 2 + 960 * 540 * 2);
  std::vector<Case> cases = {{"h2d", 1, 0}, {"d2h", 0, 1}, {"bidir", 1, 1}};
  std::vector<Variant> vars = {{"default", false, false, IN, OUT, 1}};
  if (!quick) {
    vars.push_back({"numa_affine", false, true, IN, OUT, 1});
    vars.push_back({"write_combined_src", true, false, IN, OUT, 1});
    vars.push_back({"chunk_32MB", false, false, (size_t)32 << 20, (size_t)20 << 20, 1});
    vars.push_back({"two_streams_per_dir", false, false, IN, OUT, 2});
  }
  const long cpus = sysconf(_SC_NPROCESSORS_ONLN);
  printf("{\"box\": {\"gpus\": %d, \"cpus_online\": %ld, \"in_chunk_bytes\": %zu, \"out_chunk_bytes\": %zu, \"seconds_per_case\": %.2f}}\n",
         have, cpus, IN, OUT, seconds);
  fflush(stdout);
  for (int t : {1, 8, (int)cpus}) host_memcpy_bw(t, 0.5);
  for (int n : counts) {
    if (n > have || n > 16) continue;
    Shared* sh = (Shared*)mmap(NULL, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    memset((void*)sh, 0, sizeof(Shared));
    for (int g = 0; g < n; ++g) if (fork() == 0) child(g, n + 1, sh, cases, vars, seconds, 0);
    int slot = 0;
    for (const Variant& v : vars) {
      for (const Case& c : cases) {
        barrier(sh, slot++, n + 1);
        barrier(sh, slot++, n + 1);
        double h = 0, d = 0;
        std::string per = "[";
        for (int g = 0; g < n; ++g) {
          const double hg = sh->h2d_bytes[g] / sh->secs[g] / 1e9, dg = sh->d2h_bytes[g] / sh->secs[g] / 1e9;
          h += hg; d += dg;
          char b[64]; snprintf(b, sizeof(b), "%s[%.1f, %.1f]", g ? ", " : "", hg, dg);
          per += b;
        }
        per += "]";
        printf("{\"gpus\": %d, \"case\": \"%s\", \"variant\": \"%s\", \"h2d_gbs\": %.1f, \"d2h_gbs\": %.1f, \"total_gbs\": %.1f, "
               "\"per_gpu_h2d_d2h\": %s, \"numa_node_gpu0\": %d}\n", n, c.name, v.name, h, d, h + d, per.c_str(), sh->numa_node[0]);
        fflush(stdout);
        barrier(sh, slot++, n + 1);
      }
    }
    for (int g = 0; g < n; ++g) wait(NULL);
    munmap(sh, sizeof(Shared));
  }
  return 0;
}
