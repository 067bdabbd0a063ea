"""Dict-backed context with the reference's config resolution order (the contract of the
reference's tests/utils.py DummyContext / FakeContext, tests/utils.py:323-413)."""


class Ctx:
    def __init__(self, config=None, data=None, plugins=None):
        self.config = config or {}
        self._results = {}
        for k, v in (data or {}).items():
            self._results[("run", k)] = v
        if plugins is not None:
            self._plugins = plugins

    def get_config(self, plugin, name):
        p = plugin.provides
        if p in self.config and isinstance(self.config[p], dict) and name in self.config[p]:
            return self.config[p][name]
        if f"{p}.{name}" in self.config:
            return self.config[f"{p}.{name}"]
        if name in self.config:
            return self.config[name]
        if name in getattr(plugin, "options", {}):
            return plugin.options[name].default
        return None

    def get_data(self, run_id, name):
        return self._results.get((run_id, name))

    def _set_data(self, run_id, name, data):
        self._results[(run_id, name)] = data

    def get_plugin(self, name):
        return self._plugins[name]

    def key_for(self, run_id, name):
        return f"{run_id}-{name}"
