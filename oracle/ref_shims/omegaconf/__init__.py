"""Throw-away API shim (test infrastructure): just enough of omegaconf for importing the
reference mbrl-lib hot path in the build container, where omegaconf is not installed."""


class DictConfig(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def get(self, k, default=None):
        return dict.get(self, k, default)


class ListConfig(list):
    pass


def _wrap(x):
    if isinstance(x, dict) and not isinstance(x, DictConfig):
        return DictConfig({k: _wrap(v) for k, v in x.items()})
    if isinstance(x, (list, tuple)) and not isinstance(x, ListConfig):
        return ListConfig([_wrap(v) for v in x])
    return x


class OmegaConf:
    @staticmethod
    def create(x=None):
        return _wrap({} if x is None else x)

    @staticmethod
    def to_container(x, resolve=True):
        return x

    @staticmethod
    def load(path):
        import yaml

        with open(path) as f:
            return _wrap(yaml.safe_load(f))
