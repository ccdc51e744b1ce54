// TEST INFRASTRUCTURE ONLY: esekfom.hpp includes <boost/bind.hpp> but uses nothing from it.
#pragma once
