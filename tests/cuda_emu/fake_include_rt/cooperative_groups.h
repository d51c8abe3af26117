// stand-in for <cooperative_groups.h> (whole-library build): same emulated cluster subset
#pragma once
#include "../fake_include/cooperative_groups.h"
