#pragma once
#include "../types.hpp"
