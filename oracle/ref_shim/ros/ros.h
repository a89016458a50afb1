// ROS stand-in for compiling the reference's factor sources (test infrastructure): logging and assertion macros only.
#ifndef VIML_REF_SHIM_ROS_H
#define VIML_REF_SHIM_ROS_H
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <string>
#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_DEBUG(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_INFO_STREAM(x) ((void)0)
#define ROS_WARN_STREAM(x) ((void)0)
#define ROS_DEBUG_STREAM(x) ((void)0)
#define ROS_ASSERT(c) assert(c)
#define ROS_BREAK() std::abort()
namespace ros {
class NodeHandle {};
}
#endif
