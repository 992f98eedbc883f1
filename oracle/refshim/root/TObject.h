#pragma once
#include "RootShim.h"
