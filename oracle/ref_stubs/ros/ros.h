// Stand-in for ros/ros.h: map_likelihood.h includes it for a debug publisher that
// config.h compiles out.  TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

#include <ros/console.h>
