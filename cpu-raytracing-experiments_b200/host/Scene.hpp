// Scene.hpp — kept so that `#include "Scene.hpp"` in application code still resolves: Scene and LightingAcceleration live in Renderer.hpp.
#pragma once
#include "Renderer.hpp"
