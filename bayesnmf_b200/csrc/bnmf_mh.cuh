// placeholder, filled in below
#pragma once
#include "bnmf_rng.cuh"
#include "bnmf_state.h"
