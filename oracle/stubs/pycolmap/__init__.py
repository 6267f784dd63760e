class SceneManager:
    def __init__(self, *_a, **_k):
        raise NotImplementedError('pycolmap stub (oracle harness)')
