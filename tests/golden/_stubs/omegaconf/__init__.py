"""Import-only stub (omegaconf is not installed here)."""
class OmegaConf:
    @staticmethod
    def to_container(cfg, resolve=True):
        return dict(cfg)
class DictConfig(dict):
    pass
