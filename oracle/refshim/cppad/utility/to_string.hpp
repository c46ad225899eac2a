#pragma once
#include "../cppad.hpp"
