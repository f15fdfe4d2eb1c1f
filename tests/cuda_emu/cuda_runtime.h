// stands in for <cuda_runtime.h> when the product's kernel sources are compiled for the host (tests only)
#pragma once
#include "cuda_emu.h"
