class DDPStrategy:
    def __init__(self, find_unused_parameters=False, **_):
        self.find_unused_parameters = find_unused_parameters
