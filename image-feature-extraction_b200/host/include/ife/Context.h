// One process-wide GPU context for the C++ facades, and the exception type that carries
// C-ABI failures (the reference reports failures as itk::ExceptionObject).
#ifndef IFE_B200_CONTEXT_H
#define IFE_B200_CONTEXT_H
#include <cstdlib>
#include <ostream>
#include <stdexcept>
#include <string>

#include "ife_cuda.h"

namespace ife {

class ExceptionObject : public std::runtime_error {
public:
  ExceptionObject(int code, const std::string& what) : std::runtime_error(what), m_Code(code) {}
  int GetCode() const { return m_Code; }
private:
  int m_Code;
};

inline std::ostream& operator<<(std::ostream& os, const ExceptionObject& e) { return os << e.what(); }

class CudaContext {
public:
  static CudaContext& Instance() {
    static CudaContext ctx;
    return ctx;
  }
  ife_cuda_ctx* Handle() {
    if (!m_Ctx) {
      const int rc = ife_cuda_create(m_Device, &m_Ctx);
      if (rc != IFE_OK)
        throw ExceptionObject(rc, "ife_cuda_create failed: no usable CUDA device (this library has no CPU fallback)");
      // IFE_ARITH=plain|fma selects which build of the reference the recursive Gaussian reproduces bit
      // for bit: "plain" = every multiply and add rounded (a stock x86-64 build of ITK, no -mfma),
      // "fma" = the contraction of a -mfma build (the library default, ~15 % faster)
      if (const char* a = std::getenv("IFE_ARITH")) {
        const std::string v(a);
        if (v == "plain") ife_cuda_set_arith(m_Ctx, IFE_ARITH_PLAIN);
        else if (v == "fma") ife_cuda_set_arith(m_Ctx, IFE_ARITH_FMA);
        else throw ExceptionObject(IFE_E_INVALID, "IFE_ARITH must be 'plain' or 'fma'");
      }
      // IFE_CUDA_OPTIONS="name=value,name=value": ife_cuda_set_option for each (kernel A/B switches)
      if (const char* o = std::getenv("IFE_CUDA_OPTIONS")) {
        std::string rest(o);
        while (!rest.empty()) {
          const size_t comma = rest.find(',');
          const std::string kv = rest.substr(0, comma);
          rest = comma == std::string::npos ? std::string() : rest.substr(comma + 1);
          const size_t eq = kv.find('=');
          if (kv.empty()) continue;
          if (eq == std::string::npos || ife_cuda_set_option(m_Ctx, kv.substr(0, eq).c_str(), std::atoi(kv.c_str() + eq + 1)) != IFE_OK)
            throw ExceptionObject(IFE_E_INVALID, "IFE_CUDA_OPTIONS: bad entry '" + kv + "'");
        }
      }
    }
    return m_Ctx;
  }
  void SetDevice(int device) { m_Device = device; }
  void Check(int rc) {
    if (rc != IFE_OK) throw ExceptionObject(rc, ife_cuda_last_error(m_Ctx));
  }
  ~CudaContext() { if (m_Ctx) ife_cuda_destroy(m_Ctx); }
private:
  CudaContext() {}
  ife_cuda_ctx* m_Ctx = nullptr;
  int m_Device = 0;
};

}  // namespace ife

namespace itk {
using ife::ExceptionObject;
}
#endif
