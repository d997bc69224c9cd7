// Quartet causal attention, fp32 mode.  (filled in below)
#pragma once
#include "simt_blas.cuh"
