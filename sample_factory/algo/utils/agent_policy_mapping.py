"""sample_factory.algo.utils.agent_policy_mapping (agent_policy_mapping.py:10-62): the device engine implements the
reference's sync-mode mapping -- global env index % num_policies, i.e. every policy owns a fixed 1/P of the env instances
(sample_factory_b200.multi_policy)."""


class AgentPolicyMapping:
    def __init__(self, cfg, env_info=None):
        self.num_policies = int(cfg.num_policies)

    def get_policy_for_agent(self, agent_idx: int, env_idx: int, global_env_idx: int) -> int:
        return global_env_idx % self.num_policies
