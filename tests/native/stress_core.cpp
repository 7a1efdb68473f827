// Concurrency stress of the register/launcher layer of libccm (ccm_core.cpp), meant to run
// under ThreadSanitizer: many host threads drive DIFFERENT devices at once (the contract of
// include/ccm.h), several threads contend on the SAME device, enumeration and the op trace are
// read concurrently, and whole batched transitions run side by side on disjoint device sets.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "ccm.h"

#define CHECK(x) do { int rc__ = (x); if (rc__ != 0) { char b[256]; ccm_last_error(b, sizeof b); \
  fprintf(stderr, "FAIL %s -> %d (%s) at line %d\n", #x, rc__, b, __LINE__); failures++; } } while (0)

static std::atomic<int> failures{0};

int main() {
  const int G = 16;
  CHECK(ccm_sim_topology(G, 4));
  CHECK(ccm_sim_set(-1, "reset_ms", 1));
  CHECK(ccm_sim_set(-1, "boot_ms", 2));

  // 1. one thread per device: full stage/reset/wait/query cycles
  {
    std::vector<std::thread> th;
    for (int d = 0; d < G; ++d)
      th.emplace_back([d] {
        for (int it = 0; it < 40; ++it) {
          int want = it % 3, got = -1;
          CHECK(ccm_set_cc_mode(d, want));
          CHECK(ccm_reset(d));
          CHECK(ccm_wait_for_boot(d, 1000));
          CHECK(ccm_query_cc_mode(d, &got));
          if (got != want) { fprintf(stderr, "dev %d: mode %d != %d\n", d, got, want); failures++; }
        }
      });
    // readers hammering enumeration + trace while devices change
    std::atomic<bool> stop{false};
    std::thread reader([&] {
      std::vector<ccm_dev_info> infos(64);
      std::vector<char> buf(1 << 20);
      for (int it = 0; !stop; ++it) {
        int n = 0;
        CHECK(ccm_enumerate(infos.data(), 64, &n));
        if (n != G + 4) failures++;
        if (it % 64 == 0) ccm_sim_trace(buf.data(), buf.size());   // copies + sorts the whole trace
        if (it % 1024 == 1023) ccm_sim_trace_clear();
        std::this_thread::yield();
      }
    });
    for (auto& t : th) t.join();
    stop = true;
    reader.join();
  }
  // 2. contention on ONE device: ops must serialise, never corrupt
  {
    std::vector<std::thread> th;
    for (int k = 0; k < 8; ++k)
      th.emplace_back([k] {
        for (int it = 0; it < 200; ++it) {
          int m = -1;
          int rc = ccm_query_ppcie_mode(0, &m);
          if (rc != 0 && rc != CCM_ERR_NOT_BOOTED) failures++;
          if (k == 0 && it % 10 == 0) { ccm_set_ppcie_mode(0, it / 10 % 2); ccm_reset(0); ccm_wait_for_boot(0, 1000); }
        }
      });
    for (auto& t : th) t.join();
  }
  // 3. batched transitions on disjoint halves, concurrently, plus a faulty device
  {
    CHECK(ccm_sim_set(5, "fail_op", CCM_OP_WAIT_BOOT));
    int a[8], b[8], sa[8], sb[8], ca[8], cb[8];
    for (int i = 0; i < 8; ++i) { a[i] = i; b[i] = 8 + i; }
    int rca = 0, rcb = 0;
    std::thread ta([&] { rca = ccm_transition_many(8, a, CCM_CC_ON, 0, 1000, sa, ca); });
    std::thread tb([&] { rcb = ccm_transition_many(8, b, CCM_CC_DEVTOOLS, 0, 1000, sb, cb); });
    ta.join(); tb.join();
    if (rca != CCM_ERR_FAULT || sa[5] != CCM_ERR_FAULT) { fprintf(stderr, "expected injected fault, got %d/%d\n", rca, sa[5]); failures++; }
    if (rcb != 0) { fprintf(stderr, "second batch failed: %d\n", rcb); failures++; }
    for (int i = 0; i < 8; ++i) { int m = -1; ccm_query_cc_mode(b[i], &m); if (m != CCM_CC_DEVTOOLS) failures++; }
    // scrub on a device without CUDA must fail loudly, from many threads at once
    std::vector<std::thread> th;
    for (int d = 0; d < G; ++d)
      th.emplace_back([d] { ccm_scrub_result r; if (ccm_scrub_verify(d, 1 << 20, &r) != CCM_ERR_NO_CUDA) failures++; });
    for (auto& t : th) t.join();
    ccm_scrub_result rs[16]; int all[16]; double wall = 0;
    for (int i = 0; i < G; ++i) all[i] = i;
    if (ccm_scrub_verify_many(G, all, 0, rs, &wall) != CCM_ERR_NO_CUDA) failures++;
  }
  printf("stress_core: %d failures\n", failures.load());
  return failures ? 1 : 0;
}
