#pragma once
#include "correspondence_estimation_normal_shooting.h"  // both normal-based estimators live there
