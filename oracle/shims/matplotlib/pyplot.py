def switch_backend(name):  # util/utils.py:12-14
    return None
