// model_traits.cuh — per-model switches of the replay (K2) and frontier (K3F) engines.
#pragma once
#include <type_traits>

namespace demi {

// May receive()'s operations be applied as they are issued (no staged outbox)?  Yes unless the model says otherwise
// (`static constexpr bool REPLAY_DIRECT = false`): a built-in model's receive() never exceeds its outbox, so nothing
// observable depends on staging; a loaded IR program can overflow it, and DEMI_PS_QUEUE_OVF is observable.
template <class M, class = void>
struct model_replay_direct { static constexpr bool value = true; };
template <class M>
struct model_replay_direct<M, std::void_t<decltype(M::REPLAY_DIRECT)>> { static constexpr bool value = M::REPLAY_DIRECT; };

}  // namespace demi
