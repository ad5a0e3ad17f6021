// pcl/conversions.h — toPCLPointCloud2 / fromPCLPointCloud2 live next to the blob type
#pragma once
#include "PCLPointCloud2.h"
