"""DataFeed: the {name: {"data", "attributes"}} dict env classes hand to
CUDADataManager.push_data_to_device.  API of warp_drive/utils/data_feed.py:8-104."""


class DataFeed(dict):
    def add_data(self, name, data, save_copy_and_apply_at_reset=False,
                 log_data_across_episode=False, **extra_attributes):
        attrs = {
            "save_copy_and_apply_at_reset": bool(save_copy_and_apply_at_reset),
            "log_data_across_episode": bool(log_data_across_episode),
        }
        attrs.update(extra_attributes)
        self[name] = {"data": data, "attributes": attrs}

    def add_data_list(self, data_list):
        """Entries are (name, data[, save_copy[, log]]) tuples or dicts with those keys."""
        assert isinstance(data_list, list)
        for item in data_list:
            if isinstance(item, tuple):
                assert len(item) >= 2, "name and data are strictly required"
                flags = [f for f in item[2:4]]
                save = flags[0] if len(flags) > 0 and isinstance(flags[0], bool) else False
                log = flags[1] if len(flags) > 1 and isinstance(flags[1], bool) else False
                self.add_data(item[0], item[1], save, log)
            elif isinstance(item, dict):
                self.add_data(
                    item["name"], item["data"],
                    item.get("save_copy_and_apply_at_reset", False),
                    item.get("log_data_across_episode", False),
                )
            else:
                raise TypeError("data_list entries must be tuples or dicts")

    def add_pool_for_reset(self, name, data, reset_target):
        """A pool of candidate reset values for `reset_target` (one row is drawn
        uniformly at random per done env)."""
        self.add_data(name, data, False, False, is_reset_pool=True,
                      reset_target=reset_target)
