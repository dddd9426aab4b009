// Stub: the reference kernel headers include this but the .cu files use nothing from it.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
