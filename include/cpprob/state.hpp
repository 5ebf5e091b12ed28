// cpprob-b200: the inference-mode selector of the reference API
// (/root/reference include/cpprob/state.hpp:28-33 `enum class StateType`, :35-54 `class State`).
// Only StateType::sis is served by this engine; the other enumerators are kept so that user code
// which names them still compiles, and selecting them raises instead of silently doing nothing.
#ifndef CPPROB_STATE_HPP
#define CPPROB_STATE_HPP

namespace cpprob {

enum class StateType {
    compile,
    csis,
    sis,
    dryrun
};

}  // namespace cpprob
#endif  // CPPROB_STATE_HPP
