"""Import-only stub (hydra is not installed here); the golden generator never calls these."""
def initialize(*a, **k):
    raise RuntimeError("hydra stub")
def compose(*a, **k):
    raise RuntimeError("hydra stub")
