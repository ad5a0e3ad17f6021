#pragma once
#include "registration.h"
