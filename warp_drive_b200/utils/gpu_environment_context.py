"""CUDAEnvironmentContext: what an env class must provide to run on the device.
Contract of warp_drive/utils/gpu_environment_context.py:5-45."""
import logging

from warp_drive_b200.utils.data_feed import DataFeed


class CUDAEnvironmentContext:
    def __init__(self):
        self.cuda_data_manager = None
        self.cuda_function_manager = None
        self.cuda_step = None
        self.cuda_step_function_feed = None

    def initialize_step_function_context(self, cuda_data_manager, cuda_function_manager,
                                         cuda_step_function_feed, step_function_name):
        try:
            self.cuda_data_manager = cuda_data_manager
            self.cuda_function_manager = cuda_function_manager
            cuda_function_manager.initialize_functions([step_function_name])
            self.cuda_step = cuda_function_manager.get_function(step_function_name)
            self.cuda_step_function_feed = cuda_step_function_feed
            return True
        except Exception as err:  # noqa: BLE001 - reported, caller asserts
            logging.error(err)
            return False

    def get_data_dictionary(self):
        return DataFeed()

    def get_tensor_dictionary(self):
        return DataFeed()

    def get_reset_pool_dictionary(self):
        return DataFeed()
