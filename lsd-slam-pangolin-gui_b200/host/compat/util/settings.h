// util/settings.h of the lsd-slam core: the reference's output wrappers and Keyframe.h include it but read no constant of it
#pragma once
