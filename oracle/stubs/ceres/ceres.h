// Minimal stand-in for <ceres/ceres.h> (see oracle/stubs/README.md). Test infrastructure only.
#ifndef MESHODE_STUB_CERES_CERES_
#define MESHODE_STUB_CERES_CERES_
#include "ceres/jet.h"
namespace ceres {
class CostFunction {
 public:
  virtual ~CostFunction() {}
};
// declaration-level stand-in: the reference's Create() factories are compiled but never called here
template <typename Functor, int kNumResiduals, int... Ns>
class AutoDiffCostFunction : public CostFunction {
 public:
  explicit AutoDiffCostFunction(Functor* f) : functor_(f) {}
  ~AutoDiffCostFunction() override { delete functor_; }
 private:
  Functor* functor_;
};
}  // namespace ceres
#endif
