// GemmParams and the GEMM_* / OUT_* constants live in the public header.
#pragma once
#include "../../include/morefusion_b200.h"

// voxel keys written by mf_cnn_voxelize_s2d (prev_keys): b*D^3 + flat voxel index, -1 = dropped
// point; bit 30 is set when another point of the same object falls into the same voxel (the
// consumers' fast path: a voxel with a single point needs no search for its other members)
#define MF_S2D_DUP_BIT 0x40000000
#define MF_S2D_KEY_MASK 0x3FFFFFFF
