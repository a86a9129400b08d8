"""Oracle: autoregressive rollout, restated from render_rollout.py:26-90.  TEST INFRASTRUCTURE ONLY.

CPU tensors, the oracle's own preprocess / model, the trajectory grown with torch.cat as the reference does."""
from __future__ import annotations

import torch

from . import model_ref, preprocess_ref


@torch.no_grad()
def rollout(params, data, metadata, dt, box_size, window_size, n_hidden, n_steps_mp, total_time, num_neighbors=16,
            message="sender"):
    position_traj = data["Coordinates"][:window_size].permute(1, 0, 2).float()           # :32
    temp_data = data["InternalEnergy"][:window_size]
    if temp_data.dim() == 2:
        temp_data = temp_data.unsqueeze(-1)                                               # :35-36
    temp_traj = temp_data.permute(1, 0, 2).float()                                        # :37
    for _ in range(total_time - window_size):                                             # :39
        g = preprocess_ref.preprocess(position_traj[:, -window_size:].permute(1, 0, 2),
                                      temp_traj[:, -window_size:].permute(1, 0, 2), metadata, noise_std=0.0,
                                      num_neighbors=num_neighbors, box_size=box_size, dt=dt)   # :44-52
        out = model_ref.forward(params, g["x"], g["edge_index"], g["edge_attr"], n_hidden, n_steps_mp, message)
        acc = out["acceleration"] * torch.tensor(metadata["acc_std"], dtype=torch.float32) + torch.tensor(metadata["acc_mean"], dtype=torch.float32)
        rate = out["temp_rate"] * torch.tensor(metadata["temp_rate_std"], dtype=torch.float32) + torch.tensor(metadata["temp_rate_mean"], dtype=torch.float32)
        recent_position = position_traj[:, -1]
        recent_velocity = (recent_position - position_traj[:, -2]) / dt                   # :73 (raw difference)
        new_velocity = recent_velocity + acc * dt                                         # :76
        new_position = torch.remainder(recent_position + new_velocity * dt, box_size)     # :77-80
        new_temp = temp_traj[:, -1] + rate * dt                                           # :81
        position_traj = torch.cat((position_traj, new_position.unsqueeze(1)), dim=1)      # :84
        temp_traj = torch.cat((temp_traj, new_temp.unsqueeze(1)), dim=1)                  # :85
    return {"Coordinates": position_traj.permute(1, 0, 2), "InternalEnergy": temp_traj.permute(1, 0, 2)}
