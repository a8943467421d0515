def __getattr__(name):
    raise AttributeError("matplotlib stub: pyplot.%s is not available" % name)
