// Compatibility shim: the reference header of this name, served by the GPU facade.
#pragma once
#include "../../wav_io.hpp"
