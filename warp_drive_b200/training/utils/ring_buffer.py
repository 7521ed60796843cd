"""RingBuffer / RingBufferManager: a `[capacity, ...]` batch array on the device managed as a
circular queue (API of warp_drive/training/utils/ring_buffer.py:5-87: enqueue / unroll /
isfull, manager add / get / has).  DDPG keeps its n-step windows in these: every iteration
enqueues `train_batch_size_per_env` new timesteps into arrays of `train_batch_size_per_env +
n_step - 1` slots, so the last `n_step - 1` timesteps of the previous iteration stay available
(trainer_base.py:246, trainer_ddpg.py:91-94).

The queue state is two Python integers; enqueue is one device copy into the slot and unroll
returns the time-ordered contents (a view while the queue has not wrapped, one concatenation
after)."""
import torch


class RingBuffer:
    def __init__(self, name=None, size=None, data_manager=None, tensor=None):
        self.buffer_name = f"RingBuffer_{name}"
        if tensor is None:
            assert data_manager.is_data_on_device_via_torch(name)
            tensor = data_manager.data_on_device_via_torch(name)
        self.queue = tensor
        self.size = int(tensor.shape[0]) if size is None else int(size)
        assert 0 < self.size <= tensor.shape[0], (
            f"The managed the ring buffer size could not exceed the size of the container: {name}")
        self.front = -1
        self.rear = -1
        self.current_size = 0

    def enqueue(self, data):
        assert torch.is_tensor(data)
        if self.current_size == self.size:       # full: the oldest entry is overwritten
            self.front = (self.front + 1) % self.size
            self.current_size -= 1
        if self.front == -1:
            self.front = 0
        self.rear = (self.rear + 1) % self.size
        self.queue[self.rear].copy_(data)
        self.current_size += 1

    def unroll(self):
        """Contents from the oldest to the newest entry, `[current_size, ...]`."""
        if self.current_size == 0:
            return None
        if self.rear >= self.front:
            return self.queue[self.front:self.rear + 1]
        return torch.cat((self.queue[self.front:self.size], self.queue[:self.rear + 1]), dim=0)

    def isfull(self):
        return self.current_size == self.size


class RingBufferManager(dict):
    def add(self, name, size=None, data_manager=None, tensor=None):
        self[name] = RingBuffer(name=name, size=size, data_manager=data_manager, tensor=tensor)

    def get(self, name):
        assert name in self, f"{name} is not a registered ring buffer"
        return self[name]

    def has(self, name):
        return name in self
