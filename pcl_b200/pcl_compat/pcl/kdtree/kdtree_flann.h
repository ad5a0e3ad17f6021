// pcl/kdtree/kdtree_flann.h — pcl::KdTreeFLANN<PointT> is an alias of the device-backed pcl::search::KdTree here
#pragma once
#include "../search/kdtree.h"
