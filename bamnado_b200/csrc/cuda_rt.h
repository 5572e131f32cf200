// cuda_rt.h — the CUDA runtime API for host translation units (or the test-only emulator shim, tests/emul).
#pragma once
#if defined(BND_EMUL)
#include "cuda_emul.h"
#else
#include <cuda_runtime_api.h>
#endif
