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
  // 3. two full stage -> reset -> boot sequences on disjoint halves, concurrently, plus a faulty device
  {
    CHECK(ccm_sim_set(5, "fail_op", CCM_OP_WAIT_BOOT));
    auto run_half = [](int first, int mode, int* status) {
      std::vector<std::thread> th;
      for (int i = 0; i < 8; ++i)
        th.emplace_back([=] {
          const int d = first + i;
          int rc = ccm_set_cc_mode(d, mode);
          if (rc == 0) rc = ccm_reset(d);
          if (rc == 0) rc = ccm_wait_for_boot(d, 1000);
          status[i] = rc;
        });
      for (auto& t : th) t.join();
    };
    int sa[8], sb[8];
    std::thread ta([&] { run_half(0, CCM_CC_ON, sa); });
    std::thread tb([&] { run_half(8, CCM_CC_DEVTOOLS, sb); });
    ta.join(); tb.join();
    if (sa[5] != CCM_ERR_FAULT) { fprintf(stderr, "expected injected fault, got %d\n", sa[5]); failures++; }
    for (int i = 0; i < 8; ++i) {
      if (sb[i] != 0) { fprintf(stderr, "second half failed: %d\n", sb[i]); failures++; }
      int m = -1; ccm_query_cc_mode(8 + i, &m); if (m != CCM_CC_DEVTOOLS) failures++;
    }
    // scrub on a device without CUDA must fail loudly, from many threads at once
    std::vector<std::thread> th;
    for (int d = 0; d < G; ++d)
      th.emplace_back([d] { ccm_scrub_result r; if (ccm_scrub_verify(d, 1 << 20, &r) != CCM_ERR_NO_CUDA) failures++; });
    for (auto& t : th) t.join();
    ccm_scrub_result rs[16]; int all[16]; double wall = 0;
    for (int i = 0; i < G; ++i) all[i] = i;
    if (ccm_scrub_verify_many(G, all, 0, rs, &wall) != CCM_ERR_NO_CUDA) failures++;
    // nothing is pending / held on a device without CUDA: the waits and releases are cheap no-ops
    double rel = -1, waited = -1;
    if (ccm_scrub_release_wait(0, &rel, &waited) != CCM_ERR_NO_CUDA || rel != 0 || waited != 0) failures++;
    CHECK(ccm_device_release_many(G, all, &wall));
  }
  printf("stress_core: %d failures\n", failures.load());
  return failures ? 1 : 0;
}
