// Stand-in for ros/duration.h (see ros/time.h in this directory).
#pragma once
#include <ros/time.h>
