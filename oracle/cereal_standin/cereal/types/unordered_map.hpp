// cereal stand-in (see ../cereal.hpp): test infrastructure for oracle/_ref only.
#pragma once
#include "../cereal.hpp"
