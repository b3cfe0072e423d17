// host/meep_compat/hdf5.h -- the reference main.cpp includes <hdf5.h>; what it needs is in H5Cpp.h
#include "H5Cpp.h"
