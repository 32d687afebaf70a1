// Host-side access to the same formatting / rounding code the kernels run (jmath.cuh compiled for the CPU).
#pragma once
#include <string>
#include "jmath.cuh"

namespace sgp {

const MathTables& host_math_tables();          // tables.cpp
std::string jdouble(double v);                 // java.lang.Double.toString
double host_round_sig(double v, int n);

}  // namespace sgp
