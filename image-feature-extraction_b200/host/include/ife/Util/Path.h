// Path::join with the reference's behaviour (include/ife/Util/Path.h:7-21): exactly one '/'
// between the parts, trailing separators of the first and leading ones of the second dropped.
#ifndef IFE_B200_PATH_H
#define IFE_B200_PATH_H
#include <string>

namespace ife {
namespace Path {
inline std::string join(const std::string& a, const std::string& b) {
  const size_t ea = a.find_last_not_of('/');
  std::string out = ea == std::string::npos ? std::string() : a.substr(0, ea + 1);
  out += '/';
  const size_t sb = b.find_first_not_of('/');
  if (sb != std::string::npos) out += b.substr(sb);
  return out;
}
}  // namespace Path
}  // namespace ife
#endif
