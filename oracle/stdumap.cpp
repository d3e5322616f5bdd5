// Iteration order of libstdc++'s std::unordered_map<int, T>, exposed to the Python oracle.
// The reference walks `std::unordered_map<int, FeaturePtr> features_` (src/graphbase.h:49) in GraphBase::GetFeaturesIf
// (src/graphbase.cpp:124-133); where that order selects features (gauge candidates, src/graph.cpp:291) the result depends on
// the container implementation.  The oracle mirrors the reference's insert / erase history into the same container type from the
// same standard library and reads the order back.  Test infrastructure only.
#include <unordered_map>

extern "C" {
void* um_new() { return new std::unordered_map<int, int>(); }
void um_delete(void* h) { delete static_cast<std::unordered_map<int, int>*>(h); }
void um_insert(void* h, int key) { (*static_cast<std::unordered_map<int, int>*>(h))[key] = key; }  // features_[id] = f
void um_erase(void* h, int key) { static_cast<std::unordered_map<int, int>*>(h)->erase(key); }
int um_size(void* h) { return (int)static_cast<std::unordered_map<int, int>*>(h)->size(); }
int um_keys(void* h, int* out, int max_n) {
  int n = 0;
  for (const auto& kv : *static_cast<std::unordered_map<int, int>*>(h)) {
    if (n < max_n) out[n] = kv.first;
    ++n;
  }
  return n;
}
}

// std::sort is not stable: where the reference sorts candidates whose keys tie (features initialised in the same frame have
// identical covariances), the resulting order is a property of libstdc++'s introsort and of the input order.  These helpers run
// the same std::sort on an index array in the same input order with the same comparator semantics and return the permutation.
#include <algorithm>
#include <numeric>
#include <vector>

extern "C" {
// Criteria::CandidateComparison (src/options.cpp:35-60): status descending, then score descending
void std_sort_candidates(int n, const int* status, const double* score, int* perm) {
  std::vector<int> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return (status[a] > status[b]) || (status[a] == status[b] && score[a] > score[b]); });
  std::copy(idx.begin(), idx.end(), perm);
}
// AddGroupOfFeatures' comp_fun (src/manager.cpp:486-491): count descending
void std_sort_desc(int n, const int* count, int* perm) {
  std::vector<int> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return count[a] > count[b]; });
  std::copy(idx.begin(), idx.end(), perm);
}
}

// Message reorder buffer of the reference (Estimator::MaintainBuffer, src/estimator.cpp:923-941): std::make_heap once MAX_SIZE messages
// are held, std::push_heap afterwards, front executed + std::pop_heap when more than MAX_SIZE are held; the comparator looks at the
// timestamp only (src/estimator.cpp:49-52), so the order among equal timestamps is whatever libstdc++'s heap algorithms produce.
struct MsgHeap {
  std::vector<std::pair<unsigned long long, int>> buf;  // (ts, handle)
  bool initialized = false;
  int max_size = 10;
};
static bool msg_cmp(const std::pair<unsigned long long, int>& a, const std::pair<unsigned long long, int>& b) { return a.first > b.first; }
extern "C" {
void* mh_new(int max_size) {
  auto* h = new MsgHeap();
  h->max_size = max_size;
  return h;
}
void mh_delete(void* p) { delete static_cast<MsgHeap*>(p); }
// push one message; returns 1 and the popped (ts, handle) when a message becomes due, else 0
int mh_push(void* p, unsigned long long ts, int handle, unsigned long long* out_ts, int* out_handle) {
  auto* h = static_cast<MsgHeap*>(p);
  h->buf.emplace_back(ts, handle);
  if (!h->initialized) {
    if ((int)h->buf.size() >= h->max_size) {
      std::make_heap(h->buf.begin(), h->buf.end(), msg_cmp);
      h->initialized = true;
    }
  } else {
    std::push_heap(h->buf.begin(), h->buf.end(), msg_cmp);
  }
  if (h->initialized && (int)h->buf.size() > h->max_size) {
    *out_ts = h->buf.front().first;
    *out_handle = h->buf.front().second;
    std::pop_heap(h->buf.begin(), h->buf.end(), msg_cmp);
    h->buf.pop_back();
    return 1;
  }
  return 0;
}
}
