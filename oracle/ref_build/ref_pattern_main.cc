// TEST INFRASTRUCTURE — builds against the UNMODIFIED reference sources where
// they lie under /root/reference (never copied). Drives the reference's own
// host code path for the attended-index pattern:
//   SyncMethods::Lookup + SyncMethod::operator()   (sync_methods.h:52-101, sync_methods.cc:8-117)
//   {Full,Causal,Local}AttentionPolicy::Check      (flash_attention.h:45-140)
//   {…}AttentionPolicy::IsSkipped                  (flash_attention.h:48-115)
// and prints the orders and the packed mask so that oracle/pattern.py can be
// pinned against the real thing (tests/golden/pattern_*.json).
//
// usage: ref_pattern <dims> <rule:full|causal|local> <sync_mode> <window> <log2_stride> <is_causal> <Q dims...> <K dims...>
// output (text): line1 "ref <r0> [r1]" (innermost first), line2 Q orders, line3 K orders,
//                then q lines of k chars '0'/'1'.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unordered_map>
#include "tensorflow/core/framework/tensor_shape.h"
#include "flash_attention.h"
#include "sync_methods.h"

template <int D, typename Policy>
static int run(const std::string& mode, const Policy& policy,
               const std::vector<int64_t>& qs, const std::vector<int64_t>& ks) {
  std::optional<SyncMethod<D>> sm;
  SyncMethods::Lookup<D>(mode, sm);
  if (!sm.has_value()) { fprintf(stderr, "Unsupported sync_mode: %s\n", mode.c_str()); return 2; }
  tensorflow::TensorShape Qs(qs), Ks(ks);
  const auto [ref_shape, Qmap, Kmap] = (*sm)(Qs, Ks);
  int64_t q = 1, k = 1;
  for (auto v : qs) q *= v;
  for (auto v : ks) k *= v;
  printf("ref");
  cute::for_each(ref_shape, [](auto v) { printf(" %d", int(v)); });
  printf("\n");
  std::vector<int32_t> qo(q), ko(k);
  for (int64_t i = 0; i < q; ++i) { qo[i] = cute::get<0>(Qmap(int(i))); printf(i ? " %d" : "%d", qo[i]); }
  printf("\n");
  for (int64_t i = 0; i < k; ++i) { ko[i] = cute::get<0>(Kmap(int(i))); printf(i ? " %d" : "%d", ko[i]); }
  printf("\n");
  std::string line(k, '0');
  for (int64_t i = 0; i < q; ++i) {
    for (int64_t j = 0; j < k; ++j) line[j] = policy.Check(ref_shape, qo[i], ko[j]) ? '1' : '0';
    printf("%s\n", line.c_str());
  }
  return 0;
}

template <int D>
static int dispatch(const std::string& rule, const std::string& mode, int w, int s, int c,
                    const std::vector<int64_t>& qs, const std::vector<int64_t>& ks) {
  if (rule == "full") return run<D>(mode, FullAttentionPolicy{}, qs, ks);
  if (rule == "causal") return run<D>(mode, CausalAttentionPolicy{}, qs, ks);
  if (rule == "local") return run<D>(mode, LocalAttentionPolicy(w, s, c != 0), qs, ks);
  fprintf(stderr, "bad rule\n");
  return 2;
}

int main(int argc, char** argv) {
  if (argc < 8) { fprintf(stderr, "usage: see header\n"); return 2; }
  int dims = atoi(argv[1]);
  std::string rule = argv[2], mode = argv[3];
  int w = atoi(argv[4]), s = atoi(argv[5]), c = atoi(argv[6]);
  if (argc != 7 + 2 * dims) { fprintf(stderr, "bad arg count\n"); return 2; }
  std::vector<int64_t> qs, ks;
  for (int i = 0; i < dims; ++i) qs.push_back(atoll(argv[7 + i]));
  for (int i = 0; i < dims; ++i) ks.push_back(atoll(argv[7 + dims + i]));
  if (dims == 1) return dispatch<1>(rule, mode, w, s, c, qs, ks);
  if (dims == 2) return dispatch<2>(rule, mode, w, s, c, qs, ks);
  return 2;
}
