// Host-side mirror of DenseHistogram<float> (reference include/ife/Statistics/DenseHistogram.h:
// 12-84): fixed sorted edges, bins (-inf,e0], (e0,e1], ..., (e_{n-1},inf); insert; getCounts;
// getFrequencies (counts summed as int, converted to float, divided); resetCounts;
// getNumberOfBins; operator<< (comma-separated counts).  Values are binned on the GPU
// (ife_cuda_histogram): single inserts are buffered and flushed in bulk.
#ifndef IFE_B200_DENSE_HISTOGRAM_H
#define IFE_B200_DENSE_HISTOGRAM_H
#include <cassert>
#include <cstdint>
#include <initializer_list>
#include <numeric>
#include <ostream>
#include <vector>

#include "ife/Context.h"

namespace ife {

template <typename NumType = float>
class DenseHistogram {
public:
  typedef NumType value_type;
  template <typename InputIt>
  DenseHistogram(InputIt begin, InputIt end) : m_Edges(begin, end), m_Counts(m_Edges.size() + 1) {
    assert(m_Edges.size() > 0);
  }
  DenseHistogram(std::initializer_list<value_type> edges) : m_Edges(edges), m_Counts(edges.size() + 1) {
    assert(m_Edges.size() > 0);
  }

  void insert(value_type value) {
    m_Pending.push_back((float)value);
    if (m_Pending.size() >= (1u << 22)) flush();
  }
  void insert(const float* values, size_t n) {
    flush();
    accumulate(values, n);
  }
  std::vector<value_type> getFrequencies() {
    flush();
    std::vector<value_type> f(m_Counts.size());
    value_type sum = std::accumulate(m_Counts.begin(), m_Counts.end(), 0);
    for (size_t i = 0; i < f.size(); ++i) f[i] = (value_type)m_Counts[i] / sum;
    return f;
  }
  std::vector<unsigned int> getCounts() { flush(); return m_Counts; }
  void resetCounts() { m_Pending.clear(); std::fill(m_Counts.begin(), m_Counts.end(), 0u); }
  std::size_t getNumberOfBins() const { return m_Counts.size(); }

  template <typename T>
  friend std::ostream& operator<<(std::ostream&, DenseHistogram<T>&);

private:
  void flush() {
    if (m_Pending.empty()) return;
    accumulate(m_Pending.data(), m_Pending.size());
    m_Pending.clear();
  }
  void accumulate(const float* values, size_t n) {
    std::vector<float> e(m_Edges.begin(), m_Edges.end());
    std::vector<uint32_t> c(m_Counts.size());
    CudaContext& ctx = CudaContext::Instance();
    ctx.Check(ife_cuda_histogram(ctx.Handle(), values, n, e.data(), (int)e.size(), c.data(), IFE_MEM_HOST));
    for (size_t i = 0; i < c.size(); ++i) m_Counts[i] += c[i];
  }
  std::vector<value_type> m_Edges;
  std::vector<unsigned int> m_Counts;
  std::vector<float> m_Pending;
};

template <typename T>
std::ostream& operator<<(std::ostream& os, DenseHistogram<T>& hist) {
  auto c = hist.getCounts();
  for (size_t i = 0; i < c.size(); ++i) os << (i ? "," : "") << c[i];
  return os;
}

}  // namespace ife
#endif
