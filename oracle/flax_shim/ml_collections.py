class ConfigDict(dict):
    """attribute-access dict with the two methods lwm/vqgan.py calls."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def copy_and_resolve_references(self):
        return ConfigDict(self)
