// Single translation unit for librefid_b200.so (device symbols are shared without -rdc).
#include <string.h>
#include "common.cu"
#include "tapgemm.cu"
#include "haloconv.cu"
#include "wgrad.cu"
#include "halowgrad.cu"
#include "convop.cu"
#include "elementwise.cu"
#include "engine.cu"
#include "api_test.cu"
#include "loss.cu"
#include "optim.cu"
#include "metrics.cu"
