// cvshim forwarder (test infrastructure): see ../opencv.hpp
#pragma once
#include "../opencv.hpp"
