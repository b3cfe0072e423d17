/* oracle/shim/hdf5.h -- TEST INFRASTRUCTURE: the reference main.cpp includes <hdf5.h>; everything it needs is in the H5Cpp.h recorder. */
#include "H5Cpp.h"
