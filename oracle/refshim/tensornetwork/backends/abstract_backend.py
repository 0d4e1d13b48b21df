class AbstractBackend:
    def __init__(self):
        self.name = "abstract"


BaseBackend = AbstractBackend
