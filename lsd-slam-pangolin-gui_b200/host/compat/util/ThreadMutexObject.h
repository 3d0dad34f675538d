// util/ThreadMutexObject.h of the lsd-slam core (lib/GUI.h declares members of this type): a value behind a mutex
#pragma once
#include <mutex>
template <class T> class ThreadMutexObject {
 public:
  ThreadMutexObject() {}
  explicit ThreadMutexObject(T initialValue) : object(initialValue) {}
  void assignValue(T newValue) { std::lock_guard<std::mutex> l(mutex); object = newValue; }
  void set(const T &newValue) { assignValue(newValue); }
  T getValue() { std::lock_guard<std::mutex> l(mutex); return object; }
  T &getReference() { return object; }
  std::mutex &getMutex() { return mutex; }

 private:
  T object;
  std::mutex mutex;
};
