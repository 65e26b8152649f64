// Minimal stand-in for <ATen/ATen.h> (see oracle/stubs/README.md): everything the reference's interface loops use
// comes from the torch/extension.h stand-in.  Test infrastructure only.
#ifndef MESHODE_STUB_ATEN_H_
#define MESHODE_STUB_ATEN_H_
#include <torch/extension.h>
#endif
