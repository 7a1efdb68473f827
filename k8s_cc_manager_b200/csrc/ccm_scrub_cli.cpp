// ccm-scrub — native command-line front end of the HBM scrub gate (links libccm.so).
//
//   ccm-scrub [--bytes N] [--backend sim|cudasim|sysfs] (--all | --bdf <pci address> ...)
//
// Runs the concurrent gate (ccm_scrub_verify_many) and prints ONE JSON line with the same
// fields as k8s_cc_manager_b200.devices.ScrubReport.  Exit 0: every GPU clean; 3: gate failed;
// 2: usage.  Used by the manager for CC_SCRUB_ISOLATION=process (no Python start-up in the
// child) and usable from the legacy shell engine (reference scripts/cc-manager.sh) as a
// post-reset step.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ccm.h"

static int usage(const char* argv0) {
  fprintf(stderr, "usage: %s [--bytes N] [--backend sim|cudasim|sysfs] (--all | --bdf <bdf> ...)\n", argv0);
  return 2;
}

int main(int argc, char** argv) {
  std::vector<std::string> bdfs;
  unsigned long long bytes = 0;
  bool all = false;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--all") all = true;
    else if (a == "--bdf" && i + 1 < argc) bdfs.emplace_back(argv[++i]);
    else if (a == "--bytes" && i + 1 < argc) bytes = strtoull(argv[++i], nullptr, 10);
    else if (a == "--backend" && i + 1 < argc) {
      std::string b = argv[++i];
      int be = b == "sim" ? CCM_BACKEND_SIM : b == "cudasim" ? CCM_BACKEND_CUDASIM : b == "sysfs" ? CCM_BACKEND_SYSFS : -1;
      if (be < 0 || ccm_init(be) != CCM_OK) return usage(argv[0]);
    } else return usage(argv[0]);
  }
  if (!all && bdfs.empty()) return usage(argv[0]);

  int n = 0;
  ccm_enumerate(nullptr, 0, &n);
  std::vector<ccm_dev_info> infos(n > 0 ? n : 1);
  ccm_enumerate(infos.data(), n, &n);
  std::vector<int> devs;
  std::vector<std::string> names;
  if (all) {
    for (int i = 0; i < n; ++i)
      if (infos[i].kind == CCM_KIND_GPU) { devs.push_back(infos[i].index); names.emplace_back(infos[i].bdf); }
  } else {
    for (auto& want : bdfs) {
      for (auto& c : want) c = (char)tolower(c);
      int found = -1;
      for (int i = 0; i < n; ++i)
        if (infos[i].kind == CCM_KIND_GPU && want == infos[i].bdf) found = infos[i].index;
      if (found < 0) { fprintf(stderr, "unknown GPU %s\n", want.c_str()); return 1; }
      devs.push_back(found);
      names.push_back(want);
    }
  }
  std::vector<ccm_scrub_result> res(devs.size() ? devs.size() : 1);
  double wall_ms = 0;
  ccm_scrub_verify_many((int)devs.size(), devs.data(), bytes, res.data(), &wall_ms);
  bool clean = true;
  printf("{\"wall_ms\": %.3f, \"reports\": [", wall_ms);
  for (size_t i = 0; i < devs.size(); ++i) {
    const ccm_scrub_result& r = res[i];
    clean = clean && r.status == CCM_OK && r.nonzero_bytes == 0;
    printf("%s{\"bdf\": \"%s\", \"bytes_requested\": %llu, \"bytes_scrubbed\": %llu, \"device_total_bytes\": %llu, "
           "\"nonzero_bytes\": %llu, \"ms_acquire\": %.3f, \"ms_scrub\": %.3f, \"ms_verify\": %.3f, \"ms_release\": %.3f, "
           "\"ms_total\": %.3f, \"segments\": %d, \"status\": %d, \"release_deferred\": %d, "
           "\"device_free_before\": %llu, \"bytes_unreached\": %llu, \"ms_release_wait\": %.3f, \"ms_gpu_span\": %.3f}",
           i ? ", " : "", names[i].c_str(), (unsigned long long)r.bytes_requested, (unsigned long long)r.bytes_scrubbed,
           (unsigned long long)r.device_total_bytes, (unsigned long long)r.nonzero_bytes, r.ms_acquire, r.ms_scrub,
           r.ms_verify, r.ms_release, r.ms_total, r.segments, r.status, r.release_deferred,
           (unsigned long long)r.device_free_before, (unsigned long long)r.bytes_unreached, r.ms_release_wait, r.ms_gpu_span);
  }
  printf("]}\n");
  fflush(stdout);  // the verdict is out BEFORE the contexts are torn down: a parent can act on it right away
  ccm_device_release_many((int)devs.size(), devs.data(), nullptr);  // joins the deferred HBM release, drops the contexts
  return clean ? 0 : 3;
}
