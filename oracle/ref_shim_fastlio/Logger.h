// TEST INFRASTRUCTURE ONLY: stand-in for sensor_driver/common_lib/logging/Logger.h (spdlog is absent); logging is a no-op.
#pragma once
#define LOG_INFO(...) do {} while (0)
#define LOG_WARN(...) do {} while (0)
#define LOG_ERROR(...) do {} while (0)
#define LOG_DEBUG(...) do {} while (0)
