#pragma once
#include "correspondence_rejection.h"
