class _Doc(list):
    ids = {}


def publish_doctree(source=None, **kwargs):
    return _Doc()


def publish_string(*args, **kwargs):
    return b""
