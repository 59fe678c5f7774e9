"""Import-time stand-in: the reference imports mxnet in image_iter.py:18-21 / util/utils.py:17 only for .rec/.bin loaders."""
class _Stub:
    def __getattr__(self, name):
        return _Stub()
    def __call__(self, *a, **k):
        raise RuntimeError("mxnet shim called")
ndarray = nd = io = recordio = image = _Stub()
