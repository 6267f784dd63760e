class LPIPS:
    def __init__(self, *_a, **_k):
        pass
