// oracle/shim/end_deref_guard.h — TEST INFRASTRUCTURE ONLY (force-included by oracle/Makefile.ref).
//
// The reference dereferences Workspace::end() whenever a leg's workspace holds a single plane, i.e. always outside
// rough-terrain mode: `upper_bound_it = workspace.upper_bound(h)` is end(), and `upper_bound_it->first` /
// `upper_bound_it->second` are then read and COPIED (model.cpp:530-535 Leg::getWorkplane, walk_controller.cpp:956-962
// LegStepper::calculateStanceSpanChange).  The values are not used in that case (model.cpp:537, walk_controller.cpp:971
// take the single plane directly), but copying a std::map out of the bytes that happen to follow the tree header on the
// stack is undefined behaviour: the reference's own builds survive it by luck of the stack contents, this build
// crashed.  This header makes exactly that read benign and deterministic WITHOUT touching the reference's sources: for
// the one node type involved, operator-> / operator* of libstdc++'s red-black-tree iterators answer a static, empty
// {0.0, {}} pair when the iterator is end() (recognised the way _Rb_tree_decrement recognises the header node).
// Every other iterator access is unchanged.
#ifndef SHC_SHIM_END_DEREF_GUARD_H
#define SHC_SHIM_END_DEREF_GUARD_H
#include <map>
#include <utility>

namespace shc_shim {
typedef std::pair<const double, std::map<int, double>> WorkspaceValue;
inline bool isHeader(const std::_Rb_tree_node_base* n) {
  return n->_M_color == std::_S_red && (n->_M_parent == nullptr || n->_M_parent->_M_parent == n);
}
inline WorkspaceValue* endValue() {
  static thread_local WorkspaceValue v(0.0, std::map<int, double>());
  v.second.clear();
  return &v;
}
}  // namespace shc_shim

namespace std {
template <>
inline shc_shim::WorkspaceValue* _Rb_tree_iterator<shc_shim::WorkspaceValue>::operator->() const noexcept {
  if (shc_shim::isHeader(_M_node)) return shc_shim::endValue();
  return static_cast<_Link_type>(_M_node)->_M_valptr();
}
template <>
inline shc_shim::WorkspaceValue& _Rb_tree_iterator<shc_shim::WorkspaceValue>::operator*() const noexcept {
  if (shc_shim::isHeader(_M_node)) return *shc_shim::endValue();
  return *static_cast<_Link_type>(_M_node)->_M_valptr();
}
template <>
inline const shc_shim::WorkspaceValue* _Rb_tree_const_iterator<shc_shim::WorkspaceValue>::operator->() const noexcept {
  if (shc_shim::isHeader(_M_node)) return shc_shim::endValue();
  return static_cast<_Link_type>(_M_node)->_M_valptr();
}
template <>
inline const shc_shim::WorkspaceValue& _Rb_tree_const_iterator<shc_shim::WorkspaceValue>::operator*() const noexcept {
  if (shc_shim::isHeader(_M_node)) return *shc_shim::endValue();
  return *static_cast<_Link_type>(_M_node)->_M_valptr();
}
}  // namespace std
#endif
