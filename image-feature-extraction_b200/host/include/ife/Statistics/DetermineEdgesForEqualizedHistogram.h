// Equal-frequency histogram edges from SORTED samples: same contract as the reference's
// determineEdgesForEqualizedHistogram (include/ife/Statistics/
// DetermineEdgesForEqualizedHistogram.h:21-139): nBins-1 edges are written; runs of equal
// samples move an edge to whichever end of the run is closer, the surplus / deficit of
// samples being spread over the remaining bins; std::out_of_range when there are fewer
// samples than bins.  Written over random-access ranges with indices.
#ifndef IFE_B200_DETERMINE_EDGES_H
#define IFE_B200_DETERMINE_EDGES_H
#include <algorithm>
#include <cstddef>
#include <iterator>
#include <stdexcept>

namespace ife {

template <typename InputIt, typename OutputIt>
void determineEdgesForEqualizedHistogram(InputIt first, InputIt last, OutputIt d_first, size_t nBins) {
  const auto dist = std::distance(first, last);
  if (dist < 0) throw std::logic_error("Iterator first must come before iterator last");
  const size_t n = (size_t)dist;
  if (n < nBins) throw std::out_of_range("Too many bins. Number of bins must be less or equal to number of samples");
  const size_t perBin = n / nBins;
  size_t surplus = n - perBin * nBins, deficit = 0, pos = 0;
  for (size_t edge = 0; edge + 1 < nBins; ++edge) {
    size_t step = perBin;
    const size_t left = nBins - edge;
    if (surplus) {
      const size_t share = std::max<size_t>(surplus / left, 1);
      step += share;
      surplus -= share;
    } else if (deficit) {
      const size_t share = std::max<size_t>(deficit / left, 1);
      step -= share;
      deficit -= share;
    }
    pos += step;
    const auto value = first[pos];
    const size_t runBegin = (size_t)(std::lower_bound(first, first + pos, value) - first);
    if (runBegin != pos) {  // the candidate sits inside a run of equal samples
      const size_t runEnd = (size_t)(std::upper_bound(first + pos, last, value) - first);
      if (runEnd == n) {
        pos = runBegin;
      } else {
        const size_t back = pos - runBegin, forward = runEnd - pos;
        if (back < forward || (back == forward && deficit)) {
          pos = runBegin;
          if (back > deficit) { surplus = back - deficit; deficit = 0; }
          else deficit -= back;
        } else {
          pos = runEnd;
          if (forward > surplus) { deficit = forward - surplus; surplus = 0; }
          else surplus -= forward;
        }
      }
    }
    *d_first++ = first[pos];
  }
}

}  // namespace ife
#endif
