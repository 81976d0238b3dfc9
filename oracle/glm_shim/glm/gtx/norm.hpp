// Forwards to the single-header GLM stand-in (see glm.hpp). Test infrastructure only.
#include <glm/glm.hpp>
