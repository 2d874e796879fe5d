// common.h -- the two handles AeroFLEX modules share with its GUI (reference: src/common/common_aeroflex.hpp:11-19).
// Same member names and semantics; the message queue is a small mutex-protected FIFO instead of the reference's
// lock-free spsc_queue (messages are a handful of status strings per solve).
#pragma once
#include <atomic>
#include <deque>
#include <mutex>
#include <optional>
#include <string>
#include <vector>

struct SignalHandler {
    std::atomic<bool> stop = false;
    std::atomic<bool> pause = false;
};

class MessageQueue {
    std::mutex mu_;
    std::deque<std::string> q_;
    std::size_t cap_;
public:
    explicit MessageQueue(std::size_t cap = 8) : cap_(cap) {}
    bool push(const std::string& s) {  // false when full, like spsc_queue::push
        std::lock_guard<std::mutex> l(mu_);
        if (q_.size() >= cap_) return false;
        q_.push_back(s);
        return true;
    }
    std::optional<std::string> pop() {
        std::lock_guard<std::mutex> l(mu_);
        if (q_.empty()) return std::nullopt;
        std::string s = std::move(q_.front());
        q_.pop_front();
        return s;
    }
};

struct GUIHandler {
    SignalHandler signal;
    MessageQueue msg{8};
};

namespace database {
// reference: src/common/database/database.hpp:17-31 -- the polar table the VLM viscous correction interpolates
struct airfoil {
    std::vector<double> alpha, cl, cd, cmy;
};
}  // namespace database
