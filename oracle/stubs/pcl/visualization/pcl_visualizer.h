// TEST INFRASTRUCTURE ONLY -- see pcl/point_cloud.h
#pragma once
#include <pcl/point_cloud.h>
