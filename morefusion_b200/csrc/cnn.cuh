// GemmParams and the GEMM_* / OUT_* constants live in the public header.
#pragma once
#include "../../include/morefusion_b200.h"
