// GSL shim for the reference-shim build (oracle/refshim): only what the reference's sources use.
#pragma once
#include <cmath>
#include <cstddef>
