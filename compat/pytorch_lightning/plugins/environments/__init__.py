class SLURMEnvironment:
    def __init__(self, auto_requeue=True, requeue_signal=None, **_):
        self.auto_requeue, self.requeue_signal = auto_requeue, requeue_signal
