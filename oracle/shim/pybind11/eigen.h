// TEST INFRASTRUCTURE ONLY: see pybind11.h in this directory.
#pragma once
#include "pybind11.h"
