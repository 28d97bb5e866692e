logger_initialized = {}
