// stand-in for <glm/glm.hpp> (lib/GUI.h includes it; the output wrappers use nothing of it)
#pragma once
