#pragma once
#include "transformation_estimation.h"
