#pragma once
#include "outlier_removal.h"
