// oracle/shim: see thrust/sort.h
#pragma once
#include "sort.h"
