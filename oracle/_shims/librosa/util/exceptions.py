class LibrosaError(Exception):
    pass


class ParameterError(LibrosaError):
    pass
