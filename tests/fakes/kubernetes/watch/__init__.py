class WatchScriptExhausted(BaseException):
    """Raised when the scripted watch has nothing left: ends the (infinite) watch loop.

    Derives from BaseException so neither the reference's `except ApiException` nor a
    broad `except Exception` swallows it.
    """


class Watch:
    def __init__(self):
        self._stopped = False

    def stop(self):
        self._stopped = True

    def stream(self, func, *args, **kwargs):
        from .._cluster import cluster
        c = cluster()
        c.watch_calls.append(dict(kwargs))
        if not c.watch_script:
            raise WatchScriptExhausted()
        batch = c.watch_script.pop(0)
        if isinstance(batch, BaseException):
            raise batch
        for ev in batch:
            if self._stopped:
                return
            if isinstance(ev, BaseException):
                raise ev
            if callable(ev):
                ev = ev(c)
                if ev is None:
                    continue
            yield ev
