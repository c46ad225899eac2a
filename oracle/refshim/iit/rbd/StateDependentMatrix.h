#pragma once
#include "TransformsBase.h"
