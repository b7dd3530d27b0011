// oracle/shim: forwards to the one shim header (see mxnet_shim.h)
#pragma once
#include "../mxnet_shim.h"
