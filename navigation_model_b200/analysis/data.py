"""Recorded sessions: the input format of the fitting hot path.

Host-side mirror of the reference's ``analysis/data.py`` (``:7-388``): ``Session`` (optional time-based
subsampling that skips NaNs and keeps the first non-zero reward of each window ``:39-74``, slicing
``:76-83``, orientations ``:193-200``, dict round trip ``:151-179``), ``SessionList`` (``:206-369``) and
``adjust_positions_to_maze`` (``:372-388``).  Session preparation is parameter-independent and stays on
the host; :meth:`SessionList.step_streams` / :meth:`Session.step_stream` hand the result to the device
path as flat step records.  Plotting (``:187-194``) is out of scope.
"""
import numpy as np

from .metrics import compute_orientations


class Session:
    """An experimental session: trajectory, optional reward, orientations."""

    def __init__(self, timestamps, trajectory, new_sampling_time=None, reward=None, first=None, end=None,
                 interval=None, maze=None, possible_actions=None):
        if len(trajectory) != len(timestamps):
            raise RuntimeError("Session: time and trajectory must have the same lenght")
        if reward is not None and len(reward) != len(timestamps):
            raise RuntimeError("Session: reward must have the same length as time and trajectory")
        if new_sampling_time is None:
            self._trajectory = trajectory
            self._reward = reward
            self._sampling_t = np.mean(np.diff(timestamps))
        else:
            self._sampling_t = new_sampling_time
            self._trajectory, self._reward = self._subsample(timestamps, trajectory, reward, new_sampling_time)
        if first is not None:
            self._trajectory = self._trajectory[:first]
        if end is not None:
            self._trajectory = self._trajectory[-end:]
        if interval is not None:
            self._trajectory = self._trajectory[interval[0]:interval[1]]
        self._orientations = None
        self._compute_orientations(maze=maze, possible_actions=possible_actions)

    @staticmethod
    def _subsample(timestamps, trajectory, reward, dt):
        """One sample per window of at least ``dt`` seconds: the first finite position of the window and the
        first non-zero, non-NaN reward in it (0.0 if none); windows without a finite position are dropped
        (``data.py:39-74``)."""
        traj_out = []
        rew_out = None if reward is None else []
        n = len(timestamps)
        i = 0
        while i < n - 1:
            t0 = timestamps[i]
            if np.isnan(t0):
                i += 1
                continue
            j = i + 1
            while j < n and (np.isnan(timestamps[j]) or timestamps[j] <= t0 + dt):
                j += 1
            window = np.array(trajectory[i:j])
            window = window[np.all(np.isfinite(window), axis=1)]
            if len(window) > 0:
                traj_out.append(window[0])
                if reward is not None:
                    rw = np.array(reward[i:j])
                    rw = rw[rw != 0]
                    rw = rw[~np.isnan(rw)]
                    rew_out.append(rw[0] if len(rw) > 0 else 0.0)
            i = j
        return traj_out, rew_out

    @property
    def trajectory(self):
        return self._trajectory

    @property
    def has_reward(self):
        return self._reward is not None

    @property
    def reward(self):
        return self._reward

    @property
    def orientations(self):
        return self._orientations

    @property
    def sampling_time(self):
        return self._sampling_t

    @property
    def initial_position(self):
        return self._trajectory[0]

    @property
    def initial_orientation(self):
        return self._orientations[0]

    def to_dict(self):
        d = {"sampling_time": self._sampling_t, "trajectory": self._trajectory}
        if self._reward is not None:
            d["reward"] = self._reward
        return d

    @classmethod
    def from_dict(cls, session_dict):
        s = cls.__new__(cls)
        s._trajectory = session_dict["trajectory"]
        s._sampling_t = session_dict["sampling_time"]
        s._reward = session_dict.get("reward", None)
        s._orientations = None
        s._compute_orientations()
        return s

    def _compute_orientations(self, maze=None, possible_actions=None):
        self._orientations = compute_orientations(self._trajectory, maze=maze, possible_actions=possible_actions)

    def __len__(self):
        return len(self._trajectory)

    # -- device hand-off --------------------------------------------------------------------------------
    def step_stream(self, maze, mouse=None):
        """Flat ``nm_step_rec`` records of this session (see :func:`navigation_model_b200.sweep.
        build_session_records`)."""
        from ..sweep import build_session_records
        if self._reward is None:
            raise RuntimeError("Session: a step stream needs a reward")
        return build_session_records(maze, self._trajectory, self._reward, mouse=mouse)


class SessionList:
    """An iterable list of sessions with aggregate accessors (``data.py:206-369``)."""

    def __init__(self, *args):
        if len(args) == 1 and type(args[0]) is list:
            self._sessions = args[0]
        else:
            self._sessions = [*args]
        for s in self._sessions:
            self._assert_session(s)

    def create(self, timestamps, trajectory, new_sampling_time=None, reward=None, first=None, end=None,
               interval=None, maze=None, possible_actions=None):
        self._sessions.append(Session(timestamps, trajectory, new_sampling_time, reward, first, end, interval, maze,
                                      possible_actions))

    def append(self, s):
        self._assert_session(s)
        self._sessions.append(s)

    @staticmethod
    def _concat(parts):
        out = []
        for p in parts:
            out.extend(p)  # also accepts ndarray trajectories (the reference's `list += ndarray` raises)
        return out

    @property
    def all_trajectories(self):
        return self._concat(s.trajectory for s in self._sessions)

    @property
    def all_rewards(self):
        if any(not s.has_reward for s in self._sessions):
            return None
        return self._concat(s.reward for s in self._sessions)

    @property
    def all_orientations(self):
        return self._concat(s.orientations for s in self._sessions)

    def __getitem__(self, i):
        picked = self._sessions[i]
        return SessionList(picked) if isinstance(i, slice) else picked

    def __iter__(self):
        return iter(self._sessions)

    def __len__(self):
        return len(self._sessions)

    def to_list(self):
        return [s.to_dict() for s in self._sessions]

    @classmethod
    def from_list(cls, sessions_list):
        sl = cls()
        for d in sessions_list:
            sl._sessions.append(Session.from_dict(d))
        return sl

    @staticmethod
    def _assert_session(s):
        if type(s) is not Session:
            raise RuntimeError("SessionList can only store objects of type Session")

    def __add__(self, other):
        if type(other) is not SessionList:
            raise RuntimeError("Can only add another SessionList to this SessionList")
        return SessionList(self._sessions + other._sessions)

    def __mul__(self, times):
        if type(times) is not int:
            raise RuntimeError("Can only multiply SessionList to an int")
        return SessionList(self._sessions * times)

    __rmul__ = __mul__

    def flatten(self):
        flat = Session.__new__(Session)
        flat._trajectory = self.all_trajectories
        flat._reward = self.all_rewards
        flat._sampling_t = np.mean([s.sampling_time for s in self._sessions])
        flat._orientations = self.all_orientations
        return flat

    def step_streams(self, maze, possible_actions=None):
        """One record array per session (a fresh mouse per session when ``possible_actions`` is given)."""
        from ..simulation.mouse import Mouse
        out = []
        for s in self._sessions:
            mouse = None
            if possible_actions is not None:
                mouse = Mouse(s.trajectory[0], 0.0, possible_actions, save_history=False)
            out.append(s.step_stream(maze, mouse))
        return out


def adjust_positions_to_maze(continuous_positions, maze):
    """Replace positions outside the visitable maze by the centre of the closest visitable tile; NaN
    positions are kept (``data.py:372-388``)."""
    fixed = []
    for p in continuous_positions:
        if not np.any(np.isnan(p)) and not maze.is_visitable_cont(p):
            fixed.append(maze.disc2cont(maze.get_closest_visitable_cont(p)))
        else:
            fixed.append(p)
    return fixed
