// CPU test of icpslam_b200/csrc/fiber.h (the context switch under the batched GICP host loop): N fibers with private
// stacks, each doing floating-point and libc work between a different number of yields, interleaved by a coordinator
// exactly as gicp_host.inl interleaves scans; the results must equal the same work done without fibers.
// Build twice: as is (x86-64: the assembly switch) and with -DB2_FIBER_UCONTEXT (swapcontext).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../icpslam_b200/csrc/fiber.h"

struct Job {
  FiberCtx ctx, *back = nullptr;
  std::vector<char> stack;
  int id = 0, yields = 0;
  double acc = 0;
  bool finished = false;
};
static thread_local Job* g_job = nullptr;

static double step(int id, int it, double* x) {  // callee-saved registers, SSE state and the stack all get used
  char buf[128];
  for (int k = 0; k < 6; ++k) x[k] = std::sin(x[k]) + std::sqrt(1.0 + it + 0.25 * id);
  snprintf(buf, sizeof buf, "%.17g", x[it % 6]);
  double back = 0;
  sscanf(buf, "%lf", &back);
  return back;
}
static void body(Job& j, bool with_yield) {
  double x[6];
  for (int k = 0; k < 6; ++k) x[k] = j.id + k;
  for (int it = 0; it < 40 + 7 * j.id; ++it) {
    j.acc += step(j.id, it, x);
    if (with_yield) {
      ++j.yields;
      fiber_switch(j.ctx, *j.back);
    }
  }
}
static void entry() {
  Job* j = g_job;
  body(*j, true);
  j->finished = true;
  fiber_switch(j->ctx, *j->back);
}

int main() {
  const int N = 32;
  FiberCtx main_ctx;
  std::vector<Job> jobs(N), plain(N);
  for (int i = 0; i < N; ++i) {
    jobs[i].id = plain[i].id = i;
    jobs[i].back = &main_ctx;
    jobs[i].stack.resize(256 * 1024);
    fiber_make(jobs[i].ctx, main_ctx, jobs[i].stack.data(), jobs[i].stack.size(), entry);
    body(plain[i], false);
  }
  int live = N;
  long switches = 0;
  while (live) {
    for (auto& j : jobs) {
      if (j.finished) continue;
      g_job = &j;
      fiber_switch(main_ctx, j.ctx);
      ++switches;
      if (j.finished) --live;
    }
  }
  int bad = 0;
  for (int i = 0; i < N; ++i) {
    if (std::memcmp(&jobs[i].acc, &plain[i].acc, sizeof(double)) != 0 || jobs[i].yields != 40 + 7 * i) ++bad;
  }
#ifdef B2_FIBER_ASM
  const char* kind = "asm";
#else
  const char* kind = "ucontext";
#endif
  printf("%s fibers=%d switches=%ld bad=%d\n", kind, N, switches, bad);
  return bad ? 1 : 0;
}
