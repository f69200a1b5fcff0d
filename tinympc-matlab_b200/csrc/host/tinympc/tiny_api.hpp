// Forwarding header so that reference user code keeps its `#include <tinympc/tiny_api.hpp>`
// (e.g. tinympc/TinyMPC/examples/quadrotor_hovering.cpp:22) when built against this library:
//   g++ -I tinympc-matlab_b200/csrc/host ...
#pragma once
#include "../tiny_api.hpp"
