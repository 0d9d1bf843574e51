// chamberlain_exosphere.hpp -- the header name the reference's Cython binding takes Temp_converter from
// (python/py_corona_sim.pyx:24-27: `cdef extern from "chamberlain_exosphere.hpp"`), so that the binding compiles
// unchanged against this facade.  The class lives in atmosphere.hpp.
#pragma once
#include "atmosphere.hpp"
using b200rt_host::Temp_converter;
