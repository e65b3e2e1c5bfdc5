// One process-wide GPU context for the C++ facades, and the exception type that carries
// C-ABI failures (the reference reports failures as itk::ExceptionObject).
#ifndef IFE_B200_CONTEXT_H
#define IFE_B200_CONTEXT_H
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
